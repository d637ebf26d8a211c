#!/bin/bash
# A/B of the correction-term form of the fused sweep's logits GEMM: packed bf16 pairs (default) vs two TF32 MMAs
OUT=gpurun_out; TAG=${1:-r2ba}; mkdir -p $OUT
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
for m in bf16 swap tf32; do
  echo "== budget, RMN_LGF_CORR=$m"
  RMN_LGF_CORR=$m timeout 600 python scripts/logistic_tf32x3_budget.py 2>&1 | tail -24
done
W="--workload logistic_mala --precision tf32x3 --strong --iters 1"
for rep in 1 2; do
ENVV="RMN_LGF_CORR=bf16"; one bf16_$rep $W
ENVV="RMN_LGF_CORR=tf32"; one tf32_$rep $W
done
ENVV="RMN_LGF_CORR=bf16"; one mmala_bf16 --workload logistic_mmala --precision tf32x3 --strong --iters 1
ENVV="RMN_LGF_CORR=tf32"; one mmala_tf32 --workload logistic_mmala --precision tf32x3 --strong --iters 1
timeout 900 python -m pytest tests/test_gpu_logistic.py -x -q -m gpu 2>&1 | tail -5
