#!/bin/bash
# ncu evidence for the logistic workloads (configs 4 and 5):  bash scripts/profile_logistic.sh <tag>
set -u
TAG=${1:-prof}; OUT=gpurun_out; mkdir -p $OUT
prof() {   # name  kernel-regex  skip  bench-args...
    local name=$1 regex=$2 skip=$3; shift 3
    python bench.py "$@" --no-cpu > $OUT/${TAG}_${name}_plain.log 2>&1 || { echo "plain run failed: $name"; tail -5 $OUT/${TAG}_${name}_plain.log; return; }
    ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
        --log-file $OUT/${TAG}_${name}_launches.csv python bench.py "$@" --no-cpu > $OUT/${TAG}_${name}_ncu1.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 \
        -o $OUT/${TAG}_${name}_full python bench.py "$@" --no-cpu > $OUT/${TAG}_${name}_ncu2.log 2>&1
    tail -c 300 $OUT/${TAG}_${name}_plain.log; echo
}
prof logistic lg_eval_kernel 3 --workload logistic_mala --steps 2 --warmup 3 --iters 1
prof mmala_metric lg_metric_kernel 3 --workload logistic_mmala --chains 4096 --steps 2 --warmup 3 --iters 1
prof mmala_tf32metric tf32x3_gemm_kernel 3 --workload logistic_mmala --chains 4096 --steps 2 --warmup 3 --iters 1 --precision tf32-metric
ls $OUT | grep $TAG
