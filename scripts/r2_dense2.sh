#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2i}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dense_tf32.py tests/test_gpu_logistic.py tests/test_gpu_examples.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
one() {
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
for w in 1 2 4; do ENVV="RMN_TF32_WPR=$w"; one g1000_wpr$w --workload gauss1000_mala --precision tf32x3; done
ENVV="RMN_TF32_ROWS=0"; one g1000_rows0 --workload gauss1000_mala --precision tf32x3
ENVV="A=1"; one g1000_k2048 --workload gauss1000_mala --precision tf32x3 --chains 2048
ENVV="A=1"; one g1000_f64 --workload gauss1000_mala
for w in 1 2 4; do
  RMN_TF32_WPR=$w timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches_wpr$w.csv python bench.py --workload gauss1000_mala --precision tf32x3 --steps 2 --warmup 3 --iters 5 --no-cpu > /dev/null 2>&1
  python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/${TAG}_launches_wpr$w.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > mv:
        agg[r[kn][:60]].append(float(r[mv].replace(",", "")))
print("wpr=$w", {k: "%.1f us x%d" % (sum(v) / len(v) / 1e3, len(v)) for k, v in agg.items()})
PY
done
