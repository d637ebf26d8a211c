#!/bin/bash
# split-K Fisher-metric GEMM for small chain shards: tests, then config 5 at 512 / 1,024 chains per GPU with and without it
OUT=gpurun_out; TAG=${1:-r2bz}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_logistic.py -x -q -m gpu -k "metric or mmala" 2>&1 | tail -3
one() {
  local n=$1 envv=$2; shift 2
  env $envv timeout 300 python bench.py --workload logistic_mmala --precision tf32x3 --steps 6 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g ms/step=%.4f" % (d["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f share=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"]))
except Exception as e:
    print("$n FAILED", e)
PY
}
one k512_split A=1 --strong --chains 512
one k512_nosplit RMN_MMALA_KSPLIT=1 --strong --chains 512
one k1024_split A=1 --strong --chains 1024
one k1024_nosplit RMN_MMALA_KSPLIT=1 --strong --chains 1024
