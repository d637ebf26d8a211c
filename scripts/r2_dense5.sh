#!/bin/bash
# A/B: lean finish/propose pass (48 registers, fp32 proposal arithmetic) + GEMM register cap (co-residency) + mixed tiles
OUT=gpurun_out; TAG=${1:-r2t}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_dense_tf32.py tests/test_gpu_dense_gauss.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/${TAG}_pytest.log)"
one() {
  local n=$1; shift
  env $ENVV timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-ess --no-checks "$@" > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$n.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$n", "value=%.4g e2e=%.4g ms/step=%.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), r["kernel"], "ms/launch=%.4f share=%.3f" % (r["kernel_ms_per_launch"], r["kernel_share_of_step"]), "acc=%.3f" % d["diagnostics"]["accept_rate"])
except Exception as e:
    print("$n FAILED", e)
PY
}
W="--workload gauss1000_mala --precision tf32x3"
ENVV="A=1"; one mb3 $W
ENVV="RMN_TF32_HALVES=0"; one mb3_onebranch $W
ENVV="RMN_TF32_MIXED=0"; one mb3_nomixed $W
ENVV="RMN_TF32_SPLIT_MTILES=74"; one mb3_split74 $W
ENVV="RMN_TF32_SPLIT_MTILES=37"; one mb3_split37 $W
ENVV="RIEMANN_B200_LIB=$PWD/build/lib_tc_mb2.so"; one mb2 $W
ENVV="RIEMANN_B200_LIB=$PWD/build/lib_tc_mb1.so"; one mb1 $W
ENVV="RIEMANN_B200_LIB=$PWD/build/lib_tc_mb1.so RMN_TF32_HALVES=0"; one mb1_onebranch $W
for ch in 8192 2048; do
ENVV="A=1"; one k${ch}_mb3 $W --chains $ch
ENVV="RMN_TF32_HALVES=0"; one k${ch}_mb3_onebranch $W --chains $ch
done
ENVV="A=1"; one mmala_mb3 --workload logistic_mmala --precision tf32x3 --strong --iters 1
ENVV="RIEMANN_B200_LIB=$PWD/build/lib_tc_mb1.so"; one mmala_mb1 --workload logistic_mmala --precision tf32x3 --strong --iters 1
ENVV="A=1"; one lgmala --workload logistic_mala --precision tf32x3 --iters 1
timeout 600 python -m pytest tests/test_gpu_logistic.py -x -q -m gpu > $OUT/${TAG}_pytest_lg.log 2>&1; echo "pytest lg rc=$? $(tail -1 $OUT/${TAG}_pytest_lg.log)"
