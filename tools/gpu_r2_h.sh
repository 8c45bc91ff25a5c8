#!/bin/bash
# round 2, call H: wavefront kernel (traced template, boundary prefetch), wave-sized windows
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/bench_single.py 2>&1 | tee gpurun_out/r2h_single.jsonl
B200ALIGN_TRACE=1 timeout 300 python tools/bench_e2e_trace.py 2>&1 | tail -9 | tee gpurun_out/r2h_e2e_trace.log
timeout 300 python tools/run_c2.py 1000000 3 2>&1 | tail -2
exit 0
