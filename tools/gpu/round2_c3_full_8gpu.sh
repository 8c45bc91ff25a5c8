#!/bin/bash
# C3 at specification size on 8 GPUs: 10^3 queries x 10^7 database sequences, top 100 hits per query
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/bench_c3_full.py 10000000 100 > gpurun_out/r2_c3_full_8gpu.log 2>&1
grep '"config"' gpurun_out/r2_c3_full_8gpu.log | tee gpurun_out/r2_c3_full_8gpu.json
tail -3 gpurun_out/r2_c3_full_8gpu.log | cut -c 1-300
exit 0
