#!/bin/bash
# round 2, final 8-GPU record: host-link probe, bench at N = 8 with configs (C3 database-sharded,
# gather timed)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/probe_pcie.py 2>/dev/null | grep rank | sort > gpurun_out/r2f8_probe8.txt
grep -E "rank [04]/" gpurun_out/r2f8_probe8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --c3-db 800000 > gpurun_out/r2f8_bench8.json 2> gpurun_out/r2f8_bench8.err
echo "bench8 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2f8_bench8.json',):
    d = json.load(open(f))
    print(f, round(d['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pairs_per_s'], d['parity_ok'])
    c = d.get('configs', {}).get('c3')
    if c: print('  c3', round(c['gcups_device']), round(c['gcups_wall']), c['wall_s'], c['roofline_frac'], c['parity_ok'])
PY
exit 0
