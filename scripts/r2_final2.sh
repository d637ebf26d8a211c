#!/bin/bash
# second validation of round 2 (after the single-ring fused sweep): profile of the sweep, then what the driver runs
OUT=gpurun_out; TAG=${1:-r2be}; mkdir -p $OUT
prof() {   # name kernel-regex skip -- bench args
  local n=$1 k=$2 skip=$3; shift 3
  timeout 600 python bench.py "$@" --no-cpu --no-ess --no-checks > $OUT/${TAG}_${n}_plain.json 2> $OUT/${TAG}_${n}_plain.err || { echo "$n: plain run failed"; tail -3 $OUT/${TAG}_${n}_plain.err; return; }
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_${n}_launches.csv python bench.py "$@" --no-cpu --no-ess --no-checks > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $OUT/${TAG}_${n}_full python bench.py "$@" --no-cpu --no-ess --no-checks > $OUT/${TAG}_${n}_ncu.log 2>&1
  echo "$n: $(ls -la $OUT/${TAG}_${n}_full.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
prof lgf lg_fused_sweep 3 --workload logistic_mala --precision tf32x3 --chains 1024 --iters 1 --steps 2 --warmup 3
prof mmala-sweep lg_fused_sweep 3 --workload logistic_mmala --precision tf32x3 --strong --iters 1 --steps 2 --warmup 3
RMN_LGF_TIMELINE=1 RMN_LGF_TIMELINE_FILE=$OUT/${TAG}_lgf_timeline.bin python bench.py --workload logistic_mala --precision tf32x3 --chains 1024 --iters 1 --steps 2 --warmup 3 --no-cpu --no-ess --no-checks > /dev/null 2>&1
python scripts/lgf_timeline.py $OUT/${TAG}_lgf_timeline.bin > $OUT/${TAG}_lgf_timeline.txt; tail -12 $OUT/${TAG}_lgf_timeline.txt
timeout 300 python scripts/lg_fused_check.py tiny ragged d64 config5 config4 > $OUT/${TAG}_lgf_check.txt 2>&1; tail -5 $OUT/${TAG}_lgf_check.txt
bash scripts/r2_validate.sh $TAG
