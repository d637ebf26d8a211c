#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2bq}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_changepoint.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -4
for m in 1 0 1 0; do
  RMN_CP_SLICED=$m timeout 300 python bench.py --workload changepoint --steps 10 --warmup 3 --no-cpu --no-ess --no-checks --no-configs > $OUT/${TAG}_s$m.json 2> $OUT/${TAG}_s$m.err
  python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_s$m.json").read().strip().splitlines()[-1])
print("sliced=$m value=%.4g e2e=%.4g kernel ms/launch=%.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_launch"]))
PY
done
