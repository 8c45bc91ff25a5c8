#!/bin/bash
# round 2, call u: banded kernel with the two-value ring + cp.async staging
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "band" 2>&1 | tail -4
timeout 300 python tools/bench_c4_full.py 16384 2>&1 | tail -1
timeout 300 python tools/bench_c4_full.py 100000 2>&1 | tail -1 | tee gpurun_out/r2u_c4_full.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:banded_fill -s 1 -c 1 -o gpurun_out/r2u_banded_fill \
    python tools/bench_c4_full.py 100000 > gpurun_out/r2u_ncu.log 2>&1
tail -2 gpurun_out/r2u_ncu.log
exit 0
