#!/bin/bash
# ncu evidence for precision=tf32x3 on the logistic model (config 4):  bash scripts/profile_logistic_tc.sh <tag>
set -u
TAG=${1:-prof}; OUT=gpurun_out; mkdir -p $OUT
ARGS="--workload logistic_mala --iters 1 --steps 2 --warmup 3 --no-cpu --precision tf32x3"
python bench.py $ARGS > $OUT/${TAG}_lgtc_plain.log 2>&1 || { echo plain run failed; tail -5 $OUT/${TAG}_lgtc_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_lgtc_launches.csv python bench.py $ARGS > $OUT/${TAG}_lgtc_ncu1.log 2>&1
# launches of tf32x3_gemm_kernel alternate: logits GEMM <0,3>, gradient GEMM <0,1>
ncu --set full --clock-control none --import-source on -k regex:"tf32x3_gemm_kernel<0, 3>" -s 3 -c 1 -o $OUT/${TAG}_lgtc_logits python bench.py $ARGS > $OUT/${TAG}_lgtc_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lg_tc_pointwise -s 3 -c 1 -o $OUT/${TAG}_lgtc_pointwise python bench.py $ARGS > $OUT/${TAG}_lgtc_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tf32x3_gemm_kernel<0, 1>" -s 3 -c 1 -o $OUT/${TAG}_lgtc_grad python bench.py $ARGS > $OUT/${TAG}_lgtc_ncu4.log 2>&1
ls $OUT | grep ${TAG}_lgtc
