#!/bin/bash
# 2-GPU run of the driver's bench line (multi-GPU self-checks included) + the 2-rank GPU tests
OUT=gpurun_out; TAG=${1:-r2bl}; N=${2:-2}; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("headline n_gpus=%d value=%.4g e2e=%.4g ess/s=%s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d.get("min_ess_per_sec")))
print("checks:", json.dumps(d.get("multi_gpu_checks"))[:700])
for c in d.get("configs", []):
    print("%-26s %s value=%.4g e2e=%.4g" % (c["key"], c.get("scaling"), c["value"], c["e2e"]["value"]))
PY
timeout 600 python -m pytest tests/test_gpu_row_sharded.py tests/test_gpu_pooled_adapt.py -q -m gpu 2>&1 | tail -3
