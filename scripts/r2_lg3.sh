#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2ac}; mkdir -p $OUT
one() {
  local n=$1; shift
  env $ENVV timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
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
W="--workload logistic_mala --precision tf32x3 --strong --iters 1"
for rep in 1 2; do
ENVV="A=1"; one pitch8_$rep $W
ENVV="RMN_LGF_PADX=1"; one padded_$rep $W
ENVV="RMN_LGF_PADX=2"; one pitch4_$rep $W
done
