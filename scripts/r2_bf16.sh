#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2ae}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_logistic.py -x -q -m gpu 2>&1 | tail -8
one() {
  local n=$1; shift
  env $ENVV timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g e2e=%.4g ms/step=%.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f" % r["kernel_ms_per_launch"], "acc=%.3f" % d["diagnostics"]["accept_rate"], d["clocks"]["reasons"])
except Exception as e:
    print("$n FAILED", e)
PY
}
W="--workload logistic_mmala --precision tf32x3 --strong --iters 1"
ENVV="A=1"; one bf16 $W
ENVV="RMN_MMALA_BF16=0"; one tf32 $W
ENVV="A=1"; one bf16_b $W
