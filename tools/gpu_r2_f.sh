#!/bin/bash
# round 2, call F: traceback beside the next fill
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/run_c2.py 1000000 4 2>&1 | tail -3 | tee gpurun_out/r2f_run.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['fill_ms_per_step'], d['roofline']['traceback_ms_per_step'])"
timeout 200 python tools/fuzz_parity.py 60 4000 > gpurun_out/r2f_fuzz.jsonl 2> gpurun_out/r2f_fuzz.err; echo "fuzz rc=$?"; grep -c error gpurun_out/r2f_fuzz.jsonl
exit 0
