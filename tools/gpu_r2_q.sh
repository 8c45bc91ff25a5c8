#!/bin/bash
# round 2, call q: lean intra-pair kernel - parity, full suite, single-pair latencies, ncu captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strip.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/bench_single.py > gpurun_out/r2q_single.jsonl 2>&1; cat gpurun_out/r2q_single.jsonl
timeout 100 python tools/run_single.py 1368 1345 > gpurun_out/r2q_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strip_fill -s 2 -c 1 -o gpurun_out/r2q_strip_cas9 python tools/run_single.py 1368 1345 > gpurun_out/r2q_ncu.log 2>&1
tail -2 gpurun_out/r2q_ncu.log
# the real tag fill launch (the even ones are the stand-by variant that returns at once)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:interseq_tag_fill -s 2 -c 1 -o gpurun_out/r2q_tag_fill \
    python tools/run_c2.py > gpurun_out/r2q_ncu_tag.log 2>&1
tail -2 gpurun_out/r2q_ncu_tag.log
exit 0
