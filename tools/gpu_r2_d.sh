#!/bin/bash
# round 2, call D: occupancy sensitivity of the tag kernel (profile restricted to NEFF letters)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "9 24" "10 20" "13 16" "16 13"; do
  set -- $v
  echo "== warps $1 neff $2"
  B200ALIGN_NEFF=$2 B200ALIGN_LIB=$PWD/build/libb200align_W$1.so timeout 300 python tools/run_c2.py 1000000 3 2>&1 | tail -2
done | tee gpurun_out/r2d_occupancy.log
