#!/bin/bash
# Builds build/lib_tpc.so = the product library + experiments/changepoint_tpc.cu (selected unless RMN_CP_KERNEL=lanes).
# Use with RIEMANN_B200_LIB=$PWD/build/lib_tpc.so.  Not part of __graft_entry__.build().
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
sed 's#"changepoint.cuh"#"'$PWD'/riemann_b200/csrc/changepoint.cuh"#' experiments/changepoint_tpc.cu > build/_tpc.cu
nvcc $F ${TPC_FLAGS:-} -Xptxas -v -c -o build/_tpc.o build/_tpc.cu 2>&1 | grep -E "error|Used|spill" | tail -4
nvcc $F -DRMN_WITH_TPC -c -o build/_cp_tpc.o riemann_b200/csrc/changepoint.cu
nvcc -shared -o build/lib_tpc.so build/api.o build/util.o build/small_gauss.o build/_cp_tpc.o build/_tpc.o build/dense.o build/logistic.o build/tc_gemm.o build/dense_tf32.o -lcudart 2>&1 | grep -v warning || true
ls -la build/lib_tpc.so
