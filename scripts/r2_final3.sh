#!/bin/bash
# last validation of round 2: what the driver runs, then the ncu capture of the fused sweep as shipped (fp16 correction terms)
OUT=gpurun_out; TAG=${1:-r2by}; mkdir -p $OUT
bash scripts/r2_validate.sh $TAG
prof() {   # name kernel-regex skip -- bench args
  local n=$1 k=$2 skip=$3; shift 3
  timeout 300 python bench.py "$@" --no-cpu --no-ess --no-checks > $OUT/${TAG}_${n}_plain.json 2> $OUT/${TAG}_${n}_plain.err || { echo "$n: plain run failed"; tail -3 $OUT/${TAG}_${n}_plain.err; return; }
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_${n}_launches.csv python bench.py "$@" --no-cpu --no-ess --no-checks > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $OUT/${TAG}_${n}_full python bench.py "$@" --no-cpu --no-ess --no-checks > $OUT/${TAG}_${n}_ncu.log 2>&1
  echo "$n: $(ls -la $OUT/${TAG}_${n}_full.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
prof lgf lg_fused_sweep 3 --workload logistic_mala --precision tf32x3 --iters 1 --steps 2 --warmup 3
