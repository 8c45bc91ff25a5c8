#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 100 python tools/run_single.py 152 159 | tail -1; timeout 100 python tools/run_single.py 152 159 1 | tail -1
timeout 100 python tools/run_single.py 1368 1345 | tail -1; timeout 100 python tools/run_single.py 1368 1345 1 | tail -1
timeout 100 python tools/run_single.py 1368 1345 0 50 | tail -1
timeout 100 python tools/run_single.py 30000 30000 1 500 | tail -1
timeout 300 python tools/bench_single.py 2>&1 | tail -7
exit 0
