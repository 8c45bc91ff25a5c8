#!/bin/bash
# round 2, call A: parity of the tag kernel + A/B timing against the predicate kernel
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
B200ALIGN_TAG=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_tag0.json 2> gpurun_out/r2a_bench_tag0.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_tag1.json 2> gpurun_out/r2a_bench_tag1.err
timeout 600 python tools/bench_modes.py 500000 > gpurun_out/r2a_modes.jsonl 2> gpurun_out/r2a_modes.err
timeout 200 python tools/fuzz_parity.py 120 1000 > gpurun_out/r2a_fuzz.jsonl 2> gpurun_out/r2a_fuzz.err; echo "fuzz rc=$?"
grep -c error gpurun_out/r2a_fuzz.jsonl
cat gpurun_out/r2a_bench_tag0.json gpurun_out/r2a_bench_tag1.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['fill_ms_per_step'])"
cat gpurun_out/r2a_modes.jsonl
