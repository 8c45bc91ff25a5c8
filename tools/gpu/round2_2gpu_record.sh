#!/bin/bash
# round 2, call w (2 GPUs): full GPU suite, sharded entry points under torchrun, bench at N = 1 and 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded_gpu.py 2>&1 | grep -E "sharded check|Error|error" | head
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2w_bench1.json 2> gpurun_out/r2w_bench1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --c3-db 400000 > gpurun_out/r2w_bench2.json 2> gpurun_out/r2w_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2w_bench1.json', 'gpurun_out/r2w_bench2.json'):
    d = json.load(open(f))
    print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pairs_per_s'], d['parity_ok'], d['roofline']['frac'])
    c = d.get('configs', {})
    for k in ('c3', 'c4', 'c5'):
        if k in c: print(' ', k, round(c[k]['gcups_device']), round(c[k]['gcups_wall']), c[k]['roofline_frac'], c[k]['parity_ok'])
    for k, v in c.get('latency', {}).items():
        if isinstance(v, dict): print(' ', k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ('gpu_ms', 'cpu_port_ms', 'device_ms', 'gcups_device', 'parity_ok')})
PY
exit 0
