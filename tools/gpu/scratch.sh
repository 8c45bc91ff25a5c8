#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/bench_single.py > gpurun_out/scratch_single.jsonl 2>&1; grep -E "device time|256 Cas9|max_number=1\"|score_only\"" gpurun_out/scratch_single.jsonl | cut -c 1-260
timeout 200 python tools/fuzz_parity.py 60 15000 > gpurun_out/scratch_fuzz.jsonl 2> gpurun_out/scratch_fuzz.err; echo "fuzz rc=$?"; grep -c seed gpurun_out/scratch_fuzz.jsonl; grep -c error gpurun_out/scratch_fuzz.jsonl; grep error gpurun_out/scratch_fuzz.jsonl | cut -c 1-300 | head -3
exit 0
