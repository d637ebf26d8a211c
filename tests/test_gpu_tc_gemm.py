"""
GPU: the tcgen05 / TMA / TMEM 3xTF32 product (riemann_b200/csrc/tc_gemm.cu) against an fp64
matmul.  Expected accuracy: fp32-level (|err| <~ 3e-6 relative to sum_k |a||b|), and clearly
better than a single TF32 pass (hi parts only, ~1e-3).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _split(x):
    hi = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, (x - hi).astype(np.float32)


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 256, 256), (300, 512, 1024), (1000, 1024, 1024), (64, 96, 64),
                                   (4864, 1024, 48),     # 152 tiles of 128 x 256: 148 full ones + 4 cut into 8 halves
                                   (19000, 256, 16)])    # 149 tiles, one k-block
def test_tf32x3_gemm_matches_fp64(M, N, K):
    import torch
    from riemann_b200 import _lib
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) * 3).astype(np.float32)
    Ah, Al = _split(A)
    Bh, Bl = _split(B)
    d = [torch.as_tensor(x, device="cuda") for x in (Ah, Al, Bh, Bl)]
    C = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().rmn_tf32x3_gemm(M, N, K, *[_lib.ptr(t) for t in d], _lib.ptr(C), _lib.stream_ptr()))
    torch.cuda.synchronize()
    got = C.cpu().numpy().astype(np.float64)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    err3 = np.max(np.abs(got - ref) / scale)
    err1 = np.max(np.abs(Ah.astype(np.float64) @ Bh.astype(np.float64).T - ref) / scale)
    assert np.all(np.isfinite(got))
    assert err3 < 3e-6, err3
    assert err1 > 20 * err3                      # one TF32 pass is far worse: the split matters


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (300, 2080, 4096), (70, 12, 1024)])
def test_single_pass_tf32_gemm(M, N, K):
    """rmn_tf32_gemm (PASSES = 1, 4-stage ring): TF32-level accuracy -- each operand loses its low
    13 mantissa bits (2^-11 relative), the accumulation is fp32."""
    import torch
    from riemann_b200 import _lib
    rng = np.random.default_rng(M * 7 + N + K)
    A = rng.uniform(0.0, 0.25, (M, K)).astype(np.float32)            # like the p(1-p) weights
    B = (rng.standard_normal((N, K)) * 0.1).astype(np.float32)
    dA, dB = torch.as_tensor(A, device="cuda"), torch.as_tensor(B, device="cuda")
    C = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().rmn_tf32_gemm(M, N, K, _lib.ptr(dA), _lib.ptr(dB), _lib.ptr(C), _lib.stream_ptr()))
    torch.cuda.synchronize()
    got = C.cpu().numpy().astype(np.float64)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - ref) / scale) < 1.5e-3


@pytest.mark.parametrize("M,N,K,ks", [(256, 128, 4096, 7), (1024, 100, 32 * 37, 18), (70, 64, 64, 5)])
def test_split_k_partials_sum_to_the_product(M, N, K, ks):
    """rmn_tf32x3_gemm_splitk: the contraction is cut into <= ks ranges, one partial product per range."""
    import ctypes as C
    import torch
    from riemann_b200 import _lib
    rng = np.random.default_rng(M + N + K + ks)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    Ah, Al = _split(A)
    Bh, Bl = _split(B)
    d = [torch.as_tensor(x, device="cuda") for x in (Ah, Al, Bh, Bl)]
    Cp = torch.full((ks, M, N), float("nan"), dtype=torch.float32, device="cuda")
    used = C.c_int(0)
    _lib.check(_lib.load().rmn_tf32x3_gemm_splitk(M, N, K, ks, *[_lib.ptr(t) for t in d], _lib.ptr(Cp),
                                                  C.byref(used), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert 1 <= used.value <= ks
    got = Cp[:used.value].cpu().numpy().astype(np.float64).sum(0)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - ref) / scale) < 3e-6


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 2080, 4096), (70, 12, 1024), (512, 2080, 100032)])
def test_bf16_gemm(M, N, K):
    """rmn_bf16_gemm (kind::f16, bf16 operands, fp32 accumulate): exact products of the bf16 inputs, so the result
    matches an fp64 product of the SAME bf16 values to fp32 accumulation accuracy."""
    import torch
    from riemann_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = (torch.rand((M, K), generator=g) * 0.25).to(torch.bfloat16)              # like the p(1-p) weights
    B = (torch.randn((N, K), generator=g) * 0.1).to(torch.bfloat16)
    dA, dB = A.cuda(), B.cuda()
    C = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().rmn_bf16_gemm(M, N, K, _lib.ptr(dA), _lib.ptr(dB), _lib.ptr(C), _lib.stream_ptr()))
    torch.cuda.synchronize()
    got = C.cpu().numpy().astype(np.float64)
    A64, B64 = A.to(torch.float64).numpy(), B.to(torch.float64).numpy()
    ref = A64 @ B64.T
    scale = np.abs(A64) @ np.abs(B64).T
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - ref) / scale) < 2e-6 * max(1.0, np.sqrt(K / 4096.0))
