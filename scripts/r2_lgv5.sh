#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2bv}; mkdir -p $OUT
timeout 300 python scripts/lg_fused_check.py tiny ragged d64 config5 config4 2>&1 | tail -5 | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_logistic.py tests/test_gpu_ptsampler.py -x -q -m gpu 2>&1 | tail -3
one() {
  local n=$1; shift
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g ms/step=%.4f" % (d["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f" % r["kernel_ms_per_launch"], d["clocks"])
except Exception as e:
    print("$n FAILED", e)
PY
}
one mala4 --workload logistic_mala --precision tf32x3 --strong --iters 4
