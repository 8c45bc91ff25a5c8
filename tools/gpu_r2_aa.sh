#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strip_fill -s 2 -c 1 -o gpurun_out/r2aa_strip_traced python tools/run_single.py 1368 1345 > gpurun_out/r2aa_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strip_fill -s 2 -c 1 -o gpurun_out/r2aa_strip_score python tools/run_single.py 1368 1345 1 > gpurun_out/r2aa_ncu2.log 2>&1
tail -1 gpurun_out/r2aa_ncu1.log gpurun_out/r2aa_ncu2.log
exit 0
