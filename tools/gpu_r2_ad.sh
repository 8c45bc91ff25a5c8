#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interseq_tag_fill -s 2 -c 1 -o gpurun_out/r2ad_tag_local python tools/run_mode.py 250000 1 1 0 > gpurun_out/r2ad_ncu.log 2>&1
tail -2 gpurun_out/r2ad_ncu.log
exit 0
