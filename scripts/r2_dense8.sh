#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2w}; mkdir -p $OUT
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
ENVV="A=1"; one r64 $W
ENVV="RMN_TF32_HALVES=0"; one r64_onebranch $W
ENVV="RIEMANN_B200_LIB=$PWD/build/lib_fp_mb10.so"; one r48 $W
ENVV="RIEMANN_B200_LIB=$PWD/build/lib_fp_mb10.so RMN_TF32_HALVES=0"; one r48_onebranch $W
ENVV="A=1"; one k2048_r64 $W --chains 2048
ENVV="RIEMANN_B200_LIB=$PWD/build/lib_fp_mb10.so"; one k2048_r48 $W --chains 2048
RMN_TF32_TIMELINE=$OUT/${TAG}_tl_halves.txt timeout 300 python scripts/dense_timeline.py run 16384 2>&1 | tail -8
RMN_TF32_HALVES=0 RMN_TF32_TIMELINE=$OUT/${TAG}_tl_one.txt timeout 300 python scripts/dense_timeline.py run 16384 2>&1 | tail -3
