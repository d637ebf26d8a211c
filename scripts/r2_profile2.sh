#!/bin/bash
# re-capture of the fused sweep after the MMA-issue rewrite
OUT=gpurun_out; TAG=${1:-r2q}; mkdir -p $OUT
prof() {
  local n=$1 k=$2 skip=$3; shift 3
  timeout 600 python bench.py "$@" --no-cpu --no-ess > $OUT/${TAG}_${n}_plain.json 2> $OUT/${TAG}_${n}_plain.err || { echo "$n: plain run failed"; tail -3 $OUT/${TAG}_${n}_plain.err; return; }
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_${n}_launches.csv python bench.py "$@" --no-cpu --no-ess > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $OUT/${TAG}_${n}_full python bench.py "$@" --no-cpu --no-ess > $OUT/${TAG}_${n}_ncu.log 2>&1
  echo "$n: $(ls -la $OUT/${TAG}_${n}_full.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
prof lgf lg_fused_sweep 3 --workload logistic_mala --precision tf32x3 --iters 1 --steps 2 --warmup 3
prof lgf_mmala lg_fused_sweep 3 --workload logistic_mmala --precision tf32x3 --strong --iters 1 --steps 2 --warmup 3
RMN_LGF_TIMELINE=1 RMN_LGF_TIMELINE_FILE=$OUT/${TAG}_timeline.bin python bench.py --workload logistic_mala --precision tf32x3 --iters 1 --steps 2 --warmup 3 --no-cpu --no-ess > /dev/null 2>&1
python scripts/lgf_timeline.py $OUT/${TAG}_timeline.bin | tee $OUT/${TAG}_timeline.txt
