#!/bin/bash
# round 2, 8-GPU call: bench at N = 8 with the device planner (C2 per rank, C3 database-sharded),
# the e2e timeline with 8 processes copying at once, the multi-device entry point from one process
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --c3-db 800000 > gpurun_out/r2p_bench8.json 2> gpurun_out/r2p_bench8.err
echo "bench8 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2p_bench8.json'))
print('N=8', d['value'], d['ms_per_step'], d['e2e'])
print(json.dumps(d.get('configs', {}).get('c3'))[:1200])
PY
B200ALIGN_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/bench_e2e_trace.py > gpurun_out/r2p_trace8.log 2>&1
grep -E "e2e ms" gpurun_out/r2p_trace8.log | sort | uniq -c | sort -k4 -n | tail -12
grep -E "window [0-3]:" gpurun_out/r2p_trace8.log | tail -8
timeout 300 python -m pytest tests/test_gpu_compact.py -q -k multi 2>&1 | tail -3
exit 0
