#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "local or Local or fuzz or batch" 2>&1 | tail -3
timeout 300 python tools/bench_modes.py > gpurun_out/r2ac_modes.jsonl 2>&1; grep -E "local affine|global affine, score\+trace" gpurun_out/r2ac_modes.jsonl
exit 0
