#!/bin/bash
# final ncu captures of round 2: every kernel that changed after r2p (plain run first, then launch list, then --set full)
OUT=gpurun_out; TAG=${1:-r2f}; mkdir -p $OUT
prof() {   # name kernel-regex skip extra-env -- bench args
  local n=$1 k=$2 skip=$3 envv=$4; shift 4
  env $envv timeout 600 python bench.py "$@" --no-cpu --no-ess --no-checks > $OUT/${TAG}_${n}_plain.json 2> $OUT/${TAG}_${n}_plain.err || { echo "$n: plain run failed"; tail -3 $OUT/${TAG}_${n}_plain.err; return; }
  if [ ! -f $OUT/${TAG}_${n%%-*}_launches.csv ]; then
    env $envv timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_${n%%-*}_launches.csv python bench.py "$@" --no-cpu --no-ess --no-checks > /dev/null 2>&1
  fi
  env $envv timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $OUT/${TAG}_${n}_full python bench.py "$@" --no-cpu --no-ess --no-checks > $OUT/${TAG}_${n}_ncu.log 2>&1
  echo "$n: $(ls -la $OUT/${TAG}_${n}_full.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
prof lgf lg_fused_sweep 3 A=1 --workload logistic_mala --precision tf32x3 --iters 1 --steps 2 --warmup 3
prof mmala-sweep lg_fused_sweep 3 A=1 --workload logistic_mmala --precision tf32x3 --strong --iters 1 --steps 2 --warmup 3
prof mmala-metric tf32x3_gemm 3 A=1 --workload logistic_mmala --precision tf32x3 --strong --iters 1 --steps 2 --warmup 3
prof g1000-gemm tf32x3_gemm 30 RMN_TF32_GRAPH=0 --workload gauss1000_mala --precision tf32x3 --steps 2 --warmup 3
prof g1000-pass finish_propose 30 RMN_TF32_GRAPH=0 --workload gauss1000_mala --precision tf32x3 --steps 2 --warmup 3
RMN_TF32_TIMELINE=$OUT/${TAG}_dense_timeline_raw.txt timeout 300 python scripts/dense_timeline.py run 16384 > $OUT/${TAG}_dense_timeline.txt 2>&1; tail -6 $OUT/${TAG}_dense_timeline.txt
RMN_LGF_TIMELINE=1 RMN_LGF_TIMELINE_FILE=$OUT/${TAG}_lgf_timeline.bin python bench.py --workload logistic_mala --precision tf32x3 --iters 1 --steps 2 --warmup 3 --no-cpu --no-ess --no-checks > /dev/null 2>&1
python scripts/lgf_timeline.py $OUT/${TAG}_lgf_timeline.bin > $OUT/${TAG}_lgf_timeline.txt; tail -12 $OUT/${TAG}_lgf_timeline.txt
