#!/bin/bash
# tail effect of the changepoint kernel: 592 resident blocks of 32 chains per wave (4 blocks x 148 SMs)
OUT=gpurun_out; TAG=${1:-r2bp}; mkdir -p $OUT
for K in 18944 37888 56832 65536 75776; do
  timeout 300 python bench.py --workload changepoint --chains $K --steps 8 --warmup 3 --no-cpu --no-ess --no-checks --no-configs > $OUT/${TAG}_$K.json 2> $OUT/${TAG}_$K.err
  python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_$K.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("K=$K waves=%.2f value=%.4g kernel ms/launch=%.4f  -> ms per wave-equivalent %.4f" % ($K / 18944.0, d["value"], r["kernel_ms_per_launch"], r["kernel_ms_per_launch"] / ($K / 18944.0)))
PY
done
