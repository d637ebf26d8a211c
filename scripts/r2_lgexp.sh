#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2bf}; mkdir -p $OUT
for x in 0 1 2 3; do
RMN_LGF_EXP=$x RMN_LGF_TIMELINE=1 RMN_LGF_TIMELINE_FILE=$OUT/${TAG}_tl_$x.bin python bench.py --workload logistic_mala --precision tf32x3 --chains 1024 --iters 1 --steps 2 --warmup 3 --no-cpu --no-ess --no-checks > $OUT/${TAG}_exp$x.json 2>/dev/null
echo "== exp $x"; python scripts/lgf_timeline.py $OUT/${TAG}_tl_$x.bin | head -3
python -c "
import json; d=json.loads(open('$OUT/${TAG}_exp$x.json').read().strip().splitlines()[-1]); print('ms/launch', d['roofline']['kernel_ms_per_launch'])"
done
