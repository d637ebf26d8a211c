#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2ar}; mkdir -p $OUT
one() {
  local n=$1; shift
  env $ENVV timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g e2e=%.4g ms/step=%.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f share=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"]))
except Exception as e:
    print("$n FAILED", e)
PY
}
W="--workload gauss1000_mala --precision tf32x3"
ENVV="A=1"; one default $W
ENVV="RMN_TF32_MIXED=0"; one nomixed $W
ENVV="A=1"; one default2 $W
RMN_TF32_TIMELINE=$OUT/${TAG}_tl.txt timeout 300 python scripts/dense_timeline.py run 16384 2>&1 | grep "replay\|median"
