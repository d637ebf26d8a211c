#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2j}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dense_tf32.py tests/test_gpu_dense_gauss.py tests/test_gpu_logistic.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
one() {
  local n=$1; shift
  env $ENVV timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-ess "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
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
ENVV="A=1"; one g1000_halves --workload gauss1000_mala --precision tf32x3
ENVV="RMN_TF32_HALVES=0"; one g1000_onebranch --workload gauss1000_mala --precision tf32x3
ENVV="RMN_TF32_GRAPH=0"; one g1000_nograph --workload gauss1000_mala --precision tf32x3
for ch in 8192 4096 2048; do
ENVV="A=1"; one g1000_k${ch}_halves --workload gauss1000_mala --precision tf32x3 --chains $ch
ENVV="RMN_TF32_HALVES=0"; one g1000_k${ch}_onebranch --workload gauss1000_mala --precision tf32x3 --chains $ch
done
