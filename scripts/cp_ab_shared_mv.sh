#!/bin/bash
# Round-2 A/B, one B200 (run under gpurun): the warp-shared move schedule against the product
# build of the changepoint kernel.  Variants are built here first (bash scripts/build_cp_variants.sh ...), they travel
# to the GPU box in build/.
#   V1 = -DRMN_CP_SHARED_MV=1                            the 8 chains of a warp share one move-type draw per step
#   V2 = V1 + -DRMN_CP_UNCOND_SEARCH=0                   ... and the run-boundary search only runs when a location moved
# Gates before any number counts: injected-stream parity (must be untouched: the tape carries per-chain move types) and
# the distributional gates against the reference's own chains (KS on the marginals, posterior of k) with the variant.
OUT=gpurun_out; TAG=${1:-sharedmv}; mkdir -p $OUT
D=$PWD/riemann_b200/libriemann_b200.so
V1=$PWD/build/lib_cp_mb5_SHAREDMV=1.so
V2="$PWD/build/lib_cp_mb5_SHAREDMV=1UNCONDSEARCH=0.so"
for lib in "$V1" "$V2"; do
  n=$(basename "$lib" .so)
  RIEMANN_B200_LIB="$lib" timeout 900 python -m pytest tests/test_gpu_changepoint.py tests/test_gpu_proposals.py tests/test_gpu_ks_marginals.py \
      -x -q -m gpu > $OUT/${TAG}_pytest_$n.log 2>&1
  echo "$n pytest rc=$? $(tail -1 $OUT/${TAG}_pytest_$n.log)"
done
for lib in "$D" "$V1" "$V2"; do
  n=$(basename "$lib" .so)
  RIEMANN_B200_LIB="$lib" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    print("$n", "value=%.4g" % d["value"], "acc=%.4f" % d["diagnostics"]["accept_rate"], "max_rhat=%.3f" % d["diagnostics"]["max_rhat"],
          "miness/s=%.4g" % d["min_ess_per_sec"])
except Exception as e:
    print("$n FAILED", e)
PY
done
