#!/bin/bash
# round 2, call ab: local traced tag path with in-block optimum tracking
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/fuzz_parity.py 100 11000 > gpurun_out/r2ab_fuzz.jsonl 2> gpurun_out/r2ab_fuzz.err; echo "fuzz rc=$?"; grep -c seed gpurun_out/r2ab_fuzz.jsonl; grep -c error gpurun_out/r2ab_fuzz.jsonl; grep error gpurun_out/r2ab_fuzz.jsonl | head -3
timeout 300 python tools/bench_modes.py > gpurun_out/r2ab_modes.jsonl 2>&1; grep -E "local affine|global affine, score\+trace" gpurun_out/r2ab_modes.jsonl
exit 0
