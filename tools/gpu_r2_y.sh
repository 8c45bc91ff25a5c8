#!/bin/bash
# round 2, call y: lean kernel v3 (scores x 32, flags in the values, two-LOP3 direction bytes)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strip.py -x -q 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_single.py > gpurun_out/r2y_single.jsonl 2>&1; grep -E "device time|256 Cas9|max_number=1\"|score_only\"" gpurun_out/r2y_single.jsonl
timeout 200 python tools/fuzz_parity.py 90 9000 > gpurun_out/r2y_fuzz.jsonl 2> gpurun_out/r2y_fuzz.err; echo "fuzz rc=$?"; grep -c seed gpurun_out/r2y_fuzz.jsonl; grep -c error gpurun_out/r2y_fuzz.jsonl
exit 0
