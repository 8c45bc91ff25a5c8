#!/bin/bash
# round 2, call r: lean intra-pair kernel v2 (packed handoff, 8 rows per lane for long pairs)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strip.py -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_single.py > gpurun_out/r2r_single.jsonl 2>&1; cat gpurun_out/r2r_single.jsonl
timeout 100 python tools/run_single.py 1368 1345 > gpurun_out/r2r_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strip_fill -s 2 -c 1 -o gpurun_out/r2r_strip_cas9 python tools/run_single.py 1368 1345 > gpurun_out/r2r_ncu.log 2>&1
tail -2 gpurun_out/r2r_ncu.log
exit 0
