#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2bo}; mkdir -p $OUT
RMN_LGF_TIMELINE=1 RMN_LGF_TIMELINE_FILE=$OUT/${TAG}_tl.bin python bench.py --workload logistic_mala --precision tf32x3 --chains 1024 --iters 1 --steps 2 --warmup 3 --no-cpu --no-ess --no-checks > /dev/null 2>&1
python scripts/lgf_timeline.py $OUT/${TAG}_tl.bin
