#!/bin/bash
# A/B of the changepoint kernel geometry on one B200: parity tests per lanes-per-chain setting,
# then the headline bench per (occupancy variant, lanes per chain).
OUT=gpurun_out; TAG=${1:-ab}; mkdir -p $OUT
for gl in ${GLS:-4 8 16}; do
  RMN_CP_GL=$gl timeout 900 python -m pytest tests/test_gpu_changepoint.py tests/test_gpu_proposals.py -x -q -m gpu > $OUT/${TAG}_pytest_gl$gl.log 2>&1
  echo "GL=$gl pytest rc=$? $(tail -1 $OUT/${TAG}_pytest_gl$gl.log)"
done
bench() { lib=$1; gl=$2; name=$3
  RIEMANN_B200_LIB=$lib RMN_CP_GL=$gl timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value=%.4g" % d["value"], "e2e=%.4g" % d["e2e"]["value"], "acc=%.4f" % d["diagnostics"]["accept_rate"], "miness/s=%.4g" % d["min_ess_per_sec"])
except Exception as e:
    print("$name FAILED", e)
PY
}
D=$PWD/riemann_b200/libriemann_b200.so
bench $D 4 gl4_default
bench $D 8 gl8_default
for mb in $MBS; do
  bench $PWD/build/lib_cp_mb$mb.so 4 gl4_mb$mb
  bench $PWD/build/lib_cp_mb$mb.so 8 gl8_mb$mb
done
