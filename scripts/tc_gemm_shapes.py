#!/usr/bin/env python
"""Timing of the tcgen05 GEMM variants on given shapes (CUDA events, best of 20): which part of a slow
launch is the mainloop and which the epilogue.  python scripts/tc_gemm_shapes.py"""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riemann_b200 import _lib
lib = _lib.load()

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

for (M, N, K) in ((16384, 1024, 1024), (16384, 1024, 4096), (2048, 1024, 1024), (1024, 100032, 128)):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    Cm = torch.empty(M, N, device="cuda")
    st = _lib.stream_ptr()
    t3 = timed(lambda: _lib.check(lib.rmn_tf32x3_gemm(M, N, K, _lib.ptr(A), _lib.ptr(A), _lib.ptr(B), _lib.ptr(B), _lib.ptr(Cm), st)))
    t1 = timed(lambda: _lib.check(lib.rmn_tf32_gemm(M, N, K, _lib.ptr(A), _lib.ptr(B), _lib.ptr(Cm), st)))
    torch.backends.cuda.matmul.allow_tf32 = True
    tb = timed(lambda: torch.matmul(A, B.t()))
    fl = 2.0 * M * N * K
    print("M=%d N=%d K=%d: 3-pass %.1f us (%.0f TF/s issued), 1-pass %.1f us (%.0f TF/s), cuBLAS tf32 %.1f us (%.0f TF/s)"
          % (M, N, K, t3 * 1e3, 3 * fl / t3 / 1e9, t1 * 1e3, fl / t1 / 1e9, tb * 1e3, fl / tb / 1e9))
