#!/bin/bash
# round 2: what the driver runs -- GPU suite, smoke, default bench line, reference arm
OUT=gpurun_out; TAG=${1:-r2v}; mkdir -p $OUT
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(grep -E 'passed|failed' $OUT/${TAG}_pytest.log | tail -1)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $OUT/${TAG}_smoke.log)"
( time timeout 1500 python bench.py --steps 20 --warmup 5 ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -4 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("headline value %.4g e2e %.4g min_ess/s %s cpu %.4g (%s)" % (d["value"], d["e2e"]["value"], d["min_ess_per_sec"], d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"]))
e = d["ess"]; print("ess: window/tau %.1f max_rhat %.3f seconds %.2f" % (e["window_over_tau"], e["max_rhat"], e["seconds"]))
for c in d["configs"]:
    if "error" in c: print(c); continue
    r = c["roofline"]
    print("%-24s value %.4g e2e %.4g ms/step %.3f | %s share %.3f frac %.3f | cpu %s" % (c["key"], c["value"], c["e2e"]["value"], c["ms_per_step"], r["kernel"], r["kernel_share_of_step"], r["frac"], (c["cpu_baseline"] or {}).get("value")))
PY
