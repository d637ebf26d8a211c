#!/bin/bash
# round 2: dense Gaussian tf32x3 -- row-in-registers finish/propose pass and the narrow GEMM tile; ncu of the fused logistic sweep
OUT=gpurun_out; TAG=${1:-r2g}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dense_tf32.py tests/test_gpu_tc_gemm.py tests/test_gpu_dense_gauss.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
run() {  # name, env..., chains
  local n=$1; shift
  env "$@" timeout 300 python bench.py --workload gauss1000_mala --precision tf32x3 --steps 6 --warmup 3 --no-cpu --chains $CH > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n chains=$CH", "value=%.4g e2e=%.4g ms/step=%.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f share=%.3f frac=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"], r["frac"]), "acc=%.3f" % d["diagnostics"]["accept_rate"])
except Exception as e:
    print("$n FAILED", e)
PY
}
CH=16384; run rows1 RMN_TF32_ROWS=1; run rows0 RMN_TF32_ROWS=0
CH=2048; run k2048_narrow1 RMN_TF32_NARROW=1; run k2048_narrow0 RMN_TF32_NARROW=0; run k2048_narrow0_rows0 RMN_TF32_NARROW=0 RMN_TF32_ROWS=0
# ncu: fused logistic sweep (1,024 chains) and the new finish/propose pass
ARGS="--workload logistic_mala --iters 1 --steps 2 --warmup 3 --no-cpu --precision tf32x3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_fused_sweep -s 3 -c 1 -o $OUT/${TAG}_lgfused python bench.py $ARGS > $OUT/${TAG}_ncu_lgf.log 2>&1; echo "ncu lgf rc=$?"
ARGS="--workload gauss1000_mala --precision tf32x3 --steps 2 --warmup 3 --iters 5 --no-cpu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finish_propose_rows -s 10 -c 1 -o $OUT/${TAG}_tf32rows python bench.py $ARGS > $OUT/${TAG}_ncu_rows.log 2>&1; echo "ncu rows rc=$?"
ls -la $OUT | grep ${TAG}_ | grep ncu-rep
