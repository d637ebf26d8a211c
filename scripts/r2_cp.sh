#!/bin/bash
# round 2: the restructured changepoint kernel (warp-uniform fast paths + shared move schedule)
OUT=gpurun_out; TAG=${1:-r2b}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_changepoint.py tests/test_gpu_proposals.py tests/test_gpu_ks_marginals.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
for lib in riemann_b200/libriemann_b200.so build/lib_cp_mb4.so build/lib_cp_mb6.so build/lib_cp_mb5_DIAGEVERY=16.so build/lib_cp_mb4_DIAGEVERY=16.so; do
 for sched in group chain; do
  n=$(basename $lib .so)_$sched
  RMN_CP_SCHEDULE=$sched RIEMANN_B200_LIB=$PWD/$lib timeout 300 python bench.py --workload changepoint --steps 10 --warmup 3 --no-cpu --no-ess --burn 20000 > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    print("$n", "value=%.4g e2e=%.4g" % (d["value"], d["e2e"]["value"]), "acc=%.4f" % d["diagnostics"]["accept_rate"], "ovf=%d" % d["diagnostics"]["overflows"])
except Exception as e:
    print("$n FAILED", e)
PY
 done
done
timeout 600 python scripts/cp_schedule_check.py > $OUT/${TAG}_schedule_check.json 2> $OUT/${TAG}_schedule_check.err; echo "schedule check rc=$?"
python - <<PY
import json
for r in json.load(open("$OUT/${TAG}_schedule_check.json")):
    print(r["schedule"], "z:", ["%.1f" % z for z in r["z"]], "pair:", ["%.3f" % z for z in r["neighbour_pair_corr"]], "mean k %.4f +- %.4f" % (r["pooled_mean"][1], r["pooled_mean_se"][1]), "acc %.4f" % r["accept_rate"])
PY
