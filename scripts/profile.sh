#!/bin/bash
# Profiling recipe (B200_PROFILING.md): for each workload, a plain run that must exit 0,
# then the per-launch time list, then ONE full ncu capture of the dominant kernel.
# Usage (on the GPU box):  bash scripts/profile.sh <tag>     -> gpurun_out/<tag>_*
set -u
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT
prof() {   # name  kernel-regex  skip  bench-args...
    local name=$1 regex=$2 skip=$3; shift 3
    python bench.py "$@" --no-cpu > $OUT/${TAG}_${name}_plain.log 2>&1 || { echo "plain run failed: $name"; tail -5 $OUT/${TAG}_${name}_plain.log; return; }
    ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
        --log-file $OUT/${TAG}_${name}_launches.csv python bench.py "$@" --no-cpu > $OUT/${TAG}_${name}_ncu1.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 \
        -o $OUT/${TAG}_${name}_full python bench.py "$@" --no-cpu > $OUT/${TAG}_${name}_ncu2.log 2>&1
    tail -c 600 $OUT/${TAG}_${name}_plain.log
}
prof changepoint changepoint_kernel 2 --steps 2 --warmup 3
prof gauss1000 gemm_abt_kernel 10 --workload gauss1000_mala --steps 2 --warmup 3 --iters 5
prof logistic lg_eval_kernel 3 --workload logistic_mala --steps 2 --warmup 3 --iters 1
ls -la $OUT | grep $TAG
