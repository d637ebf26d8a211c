#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2ap}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dense_tf32.py tests/test_gpu_dense_gauss.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --workload gauss1000_mala --precision tf32x3 --steps 20 --warmup 5 --no-cpu --no-ess --no-checks > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value=%.4g e2e=%.4g serial=%.4g ms/step=%.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["serial_value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f share=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"]), "acc=%.3f" % d["diagnostics"]["accept_rate"])
PY
