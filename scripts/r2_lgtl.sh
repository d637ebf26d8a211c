#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2bb}; mkdir -p $OUT
for m in bf16 tf32; do
RMN_LGF_CORR=$m RMN_LGF_TIMELINE=1 RMN_LGF_TIMELINE_FILE=$OUT/${TAG}_tl_$m.bin python bench.py --workload logistic_mala --precision tf32x3 --iters 1 --steps 2 --warmup 3 --no-cpu --no-ess --no-checks > /dev/null 2>&1
echo "== $m"; python scripts/lgf_timeline.py $OUT/${TAG}_tl_$m.bin
done
