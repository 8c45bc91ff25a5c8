#!/bin/bash
# round 2, 2-GPU call: multi-device entry point, NCCL gathers, bench at N = 2, e2e timeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_sharded_gpu.py 2>&1 | grep -E "sharded check|Error|error" | head
B200ALIGN_TRACE=1 timeout 300 python tools/bench_e2e_trace.py 2>&1 | tail -7 | tee gpurun_out/r2n_e2e_trace.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench1.json 2> gpurun_out/r2n_bench1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench2.json 2> gpurun_out/r2n_bench2.err
python -c "
import json
for f in ('gpurun_out/r2n_bench1.json','gpurun_out/r2n_bench2.json'):
    d=json.load(open(f)); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'])"
exit 0
