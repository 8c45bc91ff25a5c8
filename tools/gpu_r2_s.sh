#!/bin/bash
# round 2, call s: top-k, strip kernel with rows per lane by batch size, latencies
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_device_plan.py tests/test_gpu_strip.py -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_single.py > gpurun_out/r2s_single.jsonl 2>&1; cat gpurun_out/r2s_single.jsonl
exit 0
