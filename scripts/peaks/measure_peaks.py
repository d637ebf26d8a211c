#!/usr/bin/env python
"""
Measure the roofline denominators MEASURED_PEAKS.json (driver-written: HBM copy bandwidth, bf16
cuBLAS) does not carry, with the same method (best of 10, CUDA events, after warm-up):
fp64 DFMA, fp64 DMMA (mma.sync m8n8k4), fp32 FFMA from scripts/peaks/peak_kernels.cu; TF32 and
fp64 GEMM through cuBLAS (torch.matmul, 8192^3).  Writes profiles/measured_peaks_extra.json,
which bench.py reads for `roofline.peak` of the fp64 / tf32 kernels.

    python scripts/peaks/measure_peaks.py        (on a B200; builds build/libpeaks.so if missing)
"""
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "build", "libpeaks.so")
SRC = os.path.join(ROOT, "scripts", "peaks", "peak_kernels.cu")


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared",
                               "-Xcompiler", "-fPIC", "-o", SO, SRC])


def main():
    build()
    if "--build-only" in sys.argv:
        return
    import torch
    assert torch.cuda.is_available()
    lib = ctypes.CDLL(SO)
    lib.peak_launch.restype = ctypes.c_double
    lib.peak_launch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    blocks = sms * 8
    out = torch.empty(blocks * 256 * 2, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res = {}

    def timed(fn, reps=10):
        fn(); fn(); fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = fn(); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        return best, r

    for name, which, iters in (("fp64_dfma_tflops", 0, 4000), ("fp32_ffma_tflops", 1, 8000), ("fp64_dmma_tflops", 2, 2000)):
        t, flop = timed(lambda: lib.peak_launch(which, blocks, iters, out.data_ptr(), st))
        res[name] = flop / t / 1e12
    n = 8192
    for name, dt, tf32 in (("tf32_cublas_tflops", torch.float32, True), ("fp64_cublas_tflops", torch.float64, False),
                           ("fp32_cublas_tflops", torch.float32, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a = torch.randn(n, n, dtype=dt, device="cuda")
        b = torch.randn(n, n, dtype=dt, device="cuda")
        t, _ = timed(lambda: torch.matmul(a, b), reps=10 if dt != torch.float64 else 4)
        res[name] = 2.0 * n ** 3 / t / 1e12
        del a, b
    torch.backends.cuda.matmul.allow_tf32 = False
    q = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm", "--format=csv,noheader"],
                       capture_output=True, text=True).stdout.strip()
    res.update({"sms": sms, "gpu": q, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                "how": "best of 10 (fp64 GEMM: 4), CUDA events, after 3 warm-up launches; register-resident FMA / "
                       "mma.sync.m8n8k4.f64 loops (scripts/peaks/peak_kernels.cu, %d blocks x 256 threads); cuBLAS "
                       "GEMMs 8192^3 through torch.matmul (tf32: allow_tf32=True)" % blocks})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "measured_peaks_extra.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
