#!/bin/bash
# round 2: full GPU suite + the per-workload A/B lines
OUT=gpurun_out; TAG=${1:-r2h}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
one() {  # name, workload, extra bench args..., env via ENVV
  local n=$1; shift
  env $ENVV timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-ess "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g e2e=%.4g ms/step=%.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f share=%.3f frac=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"], r["frac"]), "acc=%.3f" % d["diagnostics"]["accept_rate"])
except Exception as e:
    print("$n FAILED", e)
PY
}
ENVV="A=1"; one cp --workload changepoint
ENVV="RMN_TF32_ROWS=1"; one g1000_rows1 --workload gauss1000_mala --precision tf32x3
ENVV="RMN_TF32_ROWS=0"; one g1000_rows0 --workload gauss1000_mala --precision tf32x3
ENVV="A=1"; one g1000_k2048 --workload gauss1000_mala --precision tf32x3 --chains 2048
ENVV="RMN_TF32_NARROW=0"; one g1000_k2048_wide --workload gauss1000_mala --precision tf32x3 --chains 2048
ENVV="A=1"; one lg_mala_8192 --workload logistic_mala --precision tf32x3 --strong
ENVV="A=1"; one lg_mala_1024 --workload logistic_mala --precision tf32x3
ENVV="A=1"; one lg_mmala_4096 --workload logistic_mmala --precision tf32x3 --strong
timeout 600 python scripts/lg_fused_check.py tiny ragged config5 config4 > $OUT/${TAG}_check.log 2>&1; echo "check rc=$?"; cat $OUT/${TAG}_check.log | cut -c1-420
