#!/bin/bash
# after fill_metric: the split-K tests, config 5 at BASELINE's 4,096 chains (unsplit path) and at 512 (split)
OUT=gpurun_out; TAG=${1:-r2cb}; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_logistic.py -x -q -m gpu -k "metric or mmala" 2>&1 | tail -2
one() {
  local n=$1; shift
  timeout 200 python bench.py --workload logistic_mmala --precision tf32x3 --steps 8 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g ms/step=%.4f" % (d["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f share=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"]), d["clocks"])
except Exception as e:
    print("$n FAILED", e)
PY
}
one k4096 --strong
one k512 --strong --chains 512
