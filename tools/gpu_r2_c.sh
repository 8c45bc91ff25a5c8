#!/bin/bash
# round 2, call C: A/B of tag-kernel variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in A B opq B4 C4; do
  echo "== variant $v"
  B200ALIGN_LIB=$PWD/build/libb200align_$v.so timeout 300 python tools/run_c2.py 1000000 3 2>&1 | tail -2
done | tee gpurun_out/r2c_variants.log
B200ALIGN_LIB=$PWD/build/libb200align_B.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compact.py -x -q 2>&1 | tail -3
