#!/bin/bash
# round 2, call t: ncu --set full of the banded fill kernel on the C4 shape
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/bench_c4_full.py 16384 > gpurun_out/r2t_c4_plain.log 2>&1; cat gpurun_out/r2t_c4_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:banded_fill -s 1 -c 1 -o gpurun_out/r2t_banded_fill \
    python tools/bench_c4_full.py 16384 > gpurun_out/r2t_ncu.log 2>&1
tail -2 gpurun_out/r2t_ncu.log
exit 0
