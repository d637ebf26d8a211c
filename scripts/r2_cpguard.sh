#!/bin/bash
# per-chain move schedules: the unguarded instantiation vs the guarded one (RMN_CP_GUARD=1), and the parity suite
OUT=gpurun_out; TAG=${1:-r2bx}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_changepoint.py -x -q -m gpu 2>&1 | tail -3
one() {
  local n=$1 envv=$2; shift 2
  env $envv timeout 300 python bench.py --workload changepoint --steps 6 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g ms/step=%.4f" % (d["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f" % r["kernel_ms_per_launch"], d["clocks"])
except Exception as e:
    print("$n FAILED", e)
PY
}
one chain_unguarded RMN_CP_SCHEDULE=chain
one chain_guarded "RMN_CP_SCHEDULE=chain RMN_CP_GUARD=1"
one group A=1
