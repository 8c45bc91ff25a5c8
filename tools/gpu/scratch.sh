#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_modes.py > gpurun_out/scratch_modes.jsonl 2>&1; cut -c 1-150 gpurun_out/scratch_modes.jsonl
timeout 300 python tools/bench_c3_search.py 200000 100000 2>&1 | tail -1 | cut -c 1-400
timeout 300 python tools/bench_c5_full.py 6000 2>&1 | tail -1 | cut -c 1-400
exit 0
