#!/bin/bash
# round 2, call o: device planner + generated batches, full GPU suite, bench with configs, ncu of the
# tag fill kernel (launch list + --set full), compute-sanitizer memcheck / racecheck on small cases
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_device_plan.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r2o_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2o_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d.get('e2e_expanded'), d.get('parity_checked'), d.get('parity_ok'))
print(d['roofline']['frac'], d['roofline']['fill_ms_per_step'])
for k, v in d.get('configs', {}).items():
    print(k, json.dumps(v)[:900])
PY
B200ALIGN_TRACE=1 timeout 300 python tools/bench_e2e_trace.py 2>&1 | tail -8 | tee gpurun_out/r2o_e2e_trace.log
# profiles (after the plain runs exited 0)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2o_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/r2o_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:interseq_tag_fill -s 3 -c 1 -o gpurun_out/r2o_tag_fill \
    python tools/run_c2.py > gpurun_out/r2o_ncu_full.log 2>&1
tail -3 gpurun_out/r2o_ncu_full.log
# sanitizer: smoke() covers packed (all modes), emitters, x-drop, single pair; plus the planner tests
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -4 gpurun_out/r2o_memcheck_smoke.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_racecheck_smoke.log 2>&1
echo "racecheck smoke rc=$?"; tail -4 gpurun_out/r2o_racecheck_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device_plan.py tests/test_gpu_compact.py -x -q -k "not 300000 and not 85249" > gpurun_out/r2o_memcheck_tests.log 2>&1
echo "memcheck tests rc=$?"; tail -4 gpurun_out/r2o_memcheck_tests.log
exit 0
