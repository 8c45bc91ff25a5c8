#!/bin/bash
# round 2, call E: persistent tag kernel (11 warps for 20-letter windows)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/run_c2.py 1000000 4 2>&1 | tail -3 | tee gpurun_out/r2e_run.log
timeout 300 python tools/bench_modes.py 500000 2>&1 | tee gpurun_out/r2e_modes.jsonl | cut -c1-160
timeout 200 python tools/fuzz_parity.py 60 3000 > gpurun_out/r2e_fuzz.jsonl 2> gpurun_out/r2e_fuzz.err; echo "fuzz rc=$?"; grep -c error gpurun_out/r2e_fuzz.jsonl
