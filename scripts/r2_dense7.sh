#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2v}; mkdir -p $OUT
W="--workload gauss1000_mala --precision tf32x3 --steps 2 --warmup 3 --no-cpu --no-ess --no-checks"
RMN_TF32_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:finish_propose -s 30 -c 1 -o $OUT/${TAG}_fp_full python bench.py $W > $OUT/${TAG}_fp_ncu.log 2>&1
echo "fp: $(ls -la $OUT/${TAG}_fp_full.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
RMN_TF32_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm -s 30 -c 1 -o $OUT/${TAG}_gemm_full python bench.py $W > $OUT/${TAG}_gemm_ncu.log 2>&1
echo "gemm: $(ls -la $OUT/${TAG}_gemm_full.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
