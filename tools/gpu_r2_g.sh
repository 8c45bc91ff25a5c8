#!/bin/bash
# round 2, call G: wavefront kernel parity + single-pair latencies; e2e timeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/bench_single.py 2>&1 | tee gpurun_out/r2g_single.jsonl
B200ALIGN_GENERAL=antidiag timeout 300 python tools/bench_single.py 2>&1 | tee gpurun_out/r2g_single_antidiag.jsonl
B200ALIGN_TRACE=1 timeout 300 python tools/bench_e2e_trace.py 2>&1 | tail -24 | tee gpurun_out/r2g_e2e_trace.log
exit 0
