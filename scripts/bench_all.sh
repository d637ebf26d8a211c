#!/bin/bash
# One bench line per BASELINE config on 1 GPU (for DESIGN.md / profiles/results table).
OUT=gpurun_out; TAG=${1:-all}; mkdir -p $OUT
run() { name=$1; shift; timeout 600 python bench.py "$@" > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err; tail -c 300 $OUT/${TAG}_$name.json | head -c 300; echo; }
run cfg2_changepoint --steps 10 --warmup 3 --cpu-seconds 8
run cfg1_gauss2d_rw --workload gauss2d_rw --steps 5 --warmup 3 --cpu-seconds 5
run cfg3_gauss1000_f64 --workload gauss1000_mala --steps 3 --warmup 3 --cpu-seconds 8
run cfg3_gauss1000_tf32x3 --workload gauss1000_mala --precision tf32x3 --steps 3 --warmup 3 --no-cpu
run cfg4_logistic_mala --workload logistic_mala --steps 2 --warmup 3 --iters 2 --cpu-seconds 10
run cfg4_logistic_mala_k8192 --workload logistic_mala --chains 8192 --steps 2 --warmup 3 --iters 1 --no-cpu
run cfg5_logistic_mmala --workload logistic_mmala --chains 4096 --steps 2 --warmup 3 --iters 1 --cpu-seconds 10
run cfg5_logistic_mmala_tf32metric --workload logistic_mmala --chains 4096 --steps 2 --warmup 3 --iters 1 --precision tf32-metric --no-cpu
run cfg4_logistic_mala_tf32x3 --workload logistic_mala --steps 2 --warmup 3 --iters 2 --precision tf32x3 --no-cpu
run cfg4_logistic_mala_k8192_tf32x3 --workload logistic_mala --chains 8192 --steps 2 --warmup 3 --iters 1 --precision tf32x3 --no-cpu
run cfg5_logistic_mmala_tf32x3 --workload logistic_mmala --chains 4096 --steps 2 --warmup 3 --iters 1 --precision tf32x3 --no-cpu
run n3_gauss2d_pt --workload gauss2d_pt --steps 5 --warmup 3 --no-cpu
