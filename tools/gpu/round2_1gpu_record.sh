#!/bin/bash
# round 2, final 1-GPU record: smoke, full GPU suite, bench with configs, launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2f1_bench.json 2> gpurun_out/r2f1_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f1_ref.json 2> gpurun_out/r2f1_ref.err; echo "ref rc=$?"; cat gpurun_out/r2f1_ref.json | head -c 600; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f1_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs none > gpurun_out/r2f1_ncu_launches.log 2>&1
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2f1_bench.json'))
print(round(d['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pairs_per_s'], d['parity_ok'], d['roofline']['frac'], d['cpu_baseline']['value'])
c = d.get('configs', {})
for k in ('c3', 'c4', 'c5'):
    print(' ', k, round(c[k]['gcups_device']), round(c[k]['gcups_wall']), c[k]['roofline_frac'], c[k]['parity_ok'])
for k, v in c.get('latency', {}).items():
    if isinstance(v, dict): print(' ', k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ('gpu_ms', 'cpu_port_ms', 'device_ms', 'gcups_device', 'parity_ok')})
PY
exit 0
