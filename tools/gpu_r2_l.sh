#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/bench_single.py 2>&1 | tee gpurun_out/r2l_single.jsonl
B200ALIGN_TRACE=1 timeout 300 python tools/bench_e2e_trace.py 2>&1 | tail -7 | tee gpurun_out/r2l_e2e_trace.log
exit 0
