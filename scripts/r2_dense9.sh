#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2aa}; mkdir -p $OUT
for lib in riemann_b200/libriemann_b200.so build/lib_fp_noldcg.so; do
 for hv in 1 0; do
  echo "== $lib halves=$hv"
  RIEMANN_B200_LIB=$PWD/$lib RMN_TF32_HALVES=$hv RMN_TF32_TIMELINE=$OUT/${TAG}_tl.txt timeout 300 python scripts/dense_timeline.py run 16384 2>&1 | grep "replay\|median"
 done
done
timeout 600 python -m pytest tests/test_gpu_dense_tf32.py -x -q -m gpu 2>&1 | tail -2
