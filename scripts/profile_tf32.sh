#!/bin/bash
# ncu evidence for the tcgen05 path (config 3, precision tf32x3):  bash scripts/profile_tf32.sh <tag>
TAG=${1:-prof}; OUT=gpurun_out; mkdir -p $OUT
ARGS="--workload gauss1000_mala --precision tf32x3 --steps 2 --warmup 3 --iters 5 --no-cpu"
python bench.py $ARGS > $OUT/${TAG}_tf32_plain.log 2>&1 || { echo plain run failed; tail -5 $OUT/${TAG}_tf32_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_tf32_launches.csv python bench.py $ARGS > $OUT/${TAG}_tf32_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm_kernel -s 10 -c 1 -o $OUT/${TAG}_tf32_full python bench.py $ARGS > $OUT/${TAG}_tf32_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:finish_propose_f32 -s 10 -c 1 -o $OUT/${TAG}_tf32fp_full python bench.py $ARGS > $OUT/${TAG}_tf32_ncu3.log 2>&1
tail -c 400 $OUT/${TAG}_tf32_plain.log; ls -la $OUT | grep ${TAG}_tf32
