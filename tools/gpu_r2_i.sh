#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 100 python tools/run_single.py 152 159 && timeout 100 python tools/run_single.py 1368 1345 > gpurun_out/r2i_plain.log 2>&1 && cat gpurun_out/r2i_plain.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavefront_fill -s 2 -c 1 -o gpurun_out/r2i_prof_wf python tools/run_single.py 1368 1345 > gpurun_out/r2i_ncu.log 2>&1
tail -3 gpurun_out/r2i_ncu.log
exit 0
