#!/bin/bash
# round 2: fused tcgen05 logistic sweep -- correctness first (short timeouts: a wrong barrier hangs), then speed
OUT=gpurun_out; TAG=${1:-r2c}; mkdir -p $OUT
timeout 180 python scripts/lg_fused_check.py tiny > $OUT/${TAG}_check_tiny.log 2>&1; echo "tiny rc=$?"; tail -3 $OUT/${TAG}_check_tiny.log
timeout 300 python scripts/lg_fused_check.py ragged d64 > $OUT/${TAG}_check_small.log 2>&1; echo "small rc=$?"; tail -3 $OUT/${TAG}_check_small.log
timeout 600 python scripts/lg_fused_check.py config5 config4 > $OUT/${TAG}_check_big.log 2>&1; echo "big rc=$?"; tail -3 $OUT/${TAG}_check_big.log
RMN_LG_FUSED=0 timeout 600 python scripts/lg_fused_check.py config5 config4 > $OUT/${TAG}_check_big_unfused.log 2>&1; echo "big unfused rc=$?"; tail -3 $OUT/${TAG}_check_big_unfused.log
timeout 900 python -m pytest tests/test_gpu_logistic.py tests/test_gpu_ks_marginals.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
for f in 1 0; do
 for wl in logistic_mala logistic_mmala; do
  RMN_LG_FUSED=$f timeout 600 python bench.py --workload $wl --precision tf32x3 --strong --steps 4 --warmup 3 --no-cpu > $OUT/${TAG}_bench_${wl}_fused$f.json 2> $OUT/${TAG}_bench_${wl}_fused$f.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_${wl}_fused$f.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$wl fused=$f", "value=%.4g e2e=%.4g ms/step=%.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.3f share=%.3f frac=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"], r["frac"]), "acc=%.3f" % d["diagnostics"]["accept_rate"])
except Exception as e:
    print("$wl fused=$f FAILED", e)
PY
 done
done
