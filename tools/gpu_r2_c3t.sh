#!/bin/bash
# 1-GPU check of the full-size C3 driver at a small database size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/bench_c3_full.py 640000 50 2>&1 | tail -3
exit 0
