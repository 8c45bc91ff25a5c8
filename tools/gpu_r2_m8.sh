#!/bin/bash
# round 2, 8-GPU call: host-link probe (H2D / D2H / both) and the bench at N = 8
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
nproc >> gpurun_out/r2m_topo.txt; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/r2m_topo.txt
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/probe_pcie.py 2>/dev/null | grep rank | sort > gpurun_out/r2m_probe8.txt
cat gpurun_out/r2m_probe8.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench8.json 2> gpurun_out/r2m_bench8.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench8.json')); print('N=8', d['value'], d['ms_per_step'], d['e2e'])"
exit 0
