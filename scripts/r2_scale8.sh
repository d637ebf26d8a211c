#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2s8}; N=${2:-8}; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("headline n_gpus=%d value=%.4g e2e=%.4g ess/s=%s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d.get("min_ess_per_sec")))
print("checks:", json.dumps(d.get("multi_gpu_checks"))[:600])
for c in d.get("configs", []):
    print("%-60s %s value=%.4g e2e=%.4g" % (c["config"]["workload"][:60], c.get("scaling"), c["value"], c["e2e"]["value"]))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err; echo "ref rc=$?"; head -c 600 $OUT/${TAG}_ref.json
