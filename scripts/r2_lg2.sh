#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2k}; mkdir -p $OUT
timeout 600 python scripts/lg_fused_check.py tiny ragged d64 config5 config4 > $OUT/${TAG}_check.log 2>&1; echo "check rc=$?"; cut -c1-330 $OUT/${TAG}_check.log
timeout 900 python -m pytest tests/test_gpu_logistic.py tests/test_gpu_proposals.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
one() {
  local n=$1; shift
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-ess "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
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
one lg_mala_8192 --workload logistic_mala --precision tf32x3 --strong
one lg_mala_1024 --workload logistic_mala --precision tf32x3
one lg_mmala_4096 --workload logistic_mmala --precision tf32x3 --strong
