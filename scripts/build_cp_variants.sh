#!/bin/bash
# Build libriemann_b200 variants that differ only in the changepoint kernel's occupancy target
# (blocks of 128 threads per SM) for A/B measurements: build/lib_cp_mb<N>.so.
# Select one with RIEMANN_B200_LIB=$PWD/build/lib_cp_mb<N>.so ; RMN_CP_GL=4|8|16 picks lanes per chain.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
# each argument: <name>[:<extra -D flags, comma separated>], e.g. 5  6:-DRMN_CP_PREFETCH=1
for spec in "$@"; do
  mb=${spec%%:*}; extra=""; [ "$spec" != "$mb" ] && extra=$(echo "${spec#*:}" | tr ',' ' ')
  tagx=$(echo "$extra" | tr -cd 'A-Z0-9=' | sed 's/DRMNCP//g'); mbn=${mb}; mb=${mb}${tagx:+_$tagx}
  nvcc $FLAGS -DRMN_CP_MINBLOCKS=$mbn $extra -c riemann_b200/csrc/changepoint.cu -o build/changepoint_mb$mb.o
  objs="build/api.o build/util.o build/small_gauss.o build/dense.o build/logistic.o build/logistic_fused.o build/tc_gemm.o build/dense_tf32.o build/comm.o build/acf.o"
  nvcc -shared -o build/lib_cp_mb$mb.so $objs build/changepoint_mb$mb.o -lcudart -ldl
  echo "built build/lib_cp_mb$mb.so"
done
