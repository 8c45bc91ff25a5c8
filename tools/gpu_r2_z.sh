#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strip.py -x -q 2>&1 | tail -3
timeout 300 python tools/bench_single.py > gpurun_out/r2z_single.jsonl 2>&1; grep -E "device time|256 Cas9|max_number=1\"|score_only\"" gpurun_out/r2z_single.jsonl
exit 0
