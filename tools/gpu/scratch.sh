#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/fuzz_parity.py 100 13050 > gpurun_out/scratch_fuzz.jsonl 2> gpurun_out/scratch_fuzz.err; echo "fuzz rc=$?"; grep -c seed gpurun_out/scratch_fuzz.jsonl; grep -c error gpurun_out/scratch_fuzz.jsonl; grep error gpurun_out/scratch_fuzz.jsonl | cut -c 1-300
exit 0
