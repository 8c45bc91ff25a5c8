#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 100 python tools/run_single.py 152 159 > gpurun_out/r2k_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront_fill -s 2 -c 1 -o gpurun_out/r2k_prof_wf_c1 python tools/run_single.py 152 159 > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log
exit 0
