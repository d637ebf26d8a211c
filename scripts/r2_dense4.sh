#!/bin/bash
# A/B of the mixed-width GEMM tile list and the cp.async-staged finish/propose pass (config 3, tf32x3)
OUT=gpurun_out; TAG=${1:-r2s}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_dense_tf32.py tests/test_gpu_dense_gauss.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
one() {
  local n=$1; shift
  env $ENVV timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g e2e=%.4g ms/step=%.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f share=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"]), "acc=%.3f" % d["diagnostics"]["accept_rate"])
except Exception as e:
    print("$n FAILED", e)
PY
}
W="--workload gauss1000_mala --precision tf32x3"
ENVV="A=1"; one new $W
ENVV="RMN_TF32_STAGED=0"; one nostaged $W
ENVV="RMN_TF32_MIXED=0"; one nomixed $W
ENVV="RMN_TF32_MIXED=0 RMN_TF32_STAGED=0"; one old $W
ENVV="RMN_TF32_HALVES=0"; one onebranch $W
ENVV="RMN_TF32_HALVES=0 RMN_TF32_STAGED=0"; one onebranch_nostaged $W
ENVV="RMN_TF32_SPLIT_MTILES=64"; one split64 $W
ENVV="RMN_TF32_SPLIT_MTILES=37"; one split37 $W
for ch in 8192 2048; do
ENVV="A=1"; one k${ch}_new $W --chains $ch
ENVV="RMN_TF32_MIXED=0 RMN_TF32_STAGED=0"; one k${ch}_old $W --chains $ch
done
