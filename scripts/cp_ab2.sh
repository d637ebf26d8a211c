#!/bin/bash
# Bench (and optionally parity-test) a list of library variants: scripts/cp_ab2.sh <tag> <gl> <lib>...
OUT=gpurun_out; TAG=$1; GL=$2; shift 2; mkdir -p $OUT
for lib in "$@"; do
  name=$(basename $lib .so)
  if [ -n "${PARITY:-}" ]; then
    RIEMANN_B200_LIB=$PWD/$lib RMN_CP_GL=$GL timeout 900 python -m pytest tests/test_gpu_changepoint.py tests/test_gpu_proposals.py -x -q -m gpu > $OUT/${TAG}_pytest_$name.log 2>&1
    echo "$name pytest rc=$? $(tail -1 $OUT/${TAG}_pytest_$name.log)"
  fi
  RIEMANN_B200_LIB=$PWD/$lib RMN_CP_GL=$GL timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$name.json").read().strip().splitlines()[-1])
    print("$name GL=$GL", "value=%.4g" % d["value"], "e2e=%.4g" % d["e2e"]["value"], "acc=%.4f" % d["diagnostics"]["accept_rate"])
except Exception as e:
    print("$name FAILED", e)
PY
done
