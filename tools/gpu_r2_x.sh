#!/bin/bash
# round 2, call x: write-combined inputs A/B at N = 1, strip kernel store addressing, probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strip.py tests/test_gpu_compact.py -x -q 2>&1 | tail -3
timeout 100 python tools/probe_pcie.py 2>&1 | grep rank
for f in "" "--wc-inputs"; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --configs none $f > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
  python -c "
import json; d=json.load(open('gpurun_out/r2x_bench.json')); print('$f', round(d['value']), d['e2e']['ms_per_step'], d['parity_ok'])"
done
timeout 300 python tools/bench_single.py 2>&1 | grep -E "device time|256 Cas9|max_number=1\"|score_only\"" 
exit 0
