#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2bd}; mkdir -p $OUT
one() {
  local n=$1; shift
  env $ENVV timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
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
for ns in 37 23 18 9; do ENVV="RMN_LGF_NS=$ns"; one mmala_ns$ns --workload logistic_mmala --precision tf32x3 --strong --iters 4; done
for ns in 37 23 16 9; do ENVV="RMN_LGF_NS=$ns"; one mala_ns$ns --workload logistic_mala --precision tf32x3 --strong --iters 2; done
ENVV="RMN_LGF_TIMELINE=1 RMN_LGF_TIMELINE_FILE=$OUT/${TAG}_tl5.bin"; one tl5 --workload logistic_mmala --precision tf32x3 --strong --iters 1 --steps 2
python - <<PY
import numpy as np
t = np.fromfile("$OUT/${TAG}_tl5.bin", dtype=np.int64).reshape(256, 8)
ok = np.all(t > 0, axis=1); t = t[ok]
print("config 5 tiles stamped", len(t), "median cycles per tile", np.median(np.diff(t[:,1])), "first->last g1_issued", t[-1,1]-t[0,1], "pw math", np.median(t[:,6]-t[:,5]), "st..arrive", np.median(t[:,7]-t[:,6]))
PY
