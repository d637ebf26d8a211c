#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2z}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_changepoint.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --no-cpu --no-ess --no-checks > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("headline value=%.4g e2e=%.4g serial=%.4g" % (d["value"], d["e2e"]["value"], d["e2e"]["serial_value"]))
for c in d.get("configs", []):
    print("%-70s value=%.4g e2e=%.4g serial=%.4g" % (c["config"]["workload"][:70], c["value"], c["e2e"]["value"], c["e2e"]["serial_value"]))
PY
