"""GPU: the device Philox4x32-10 equals the Random123 KATs and the numpy restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_philox_kat_and_matches_numpy():
    import torch
    from riemann_b200 import _lib
    from oracle import philox
    ctr = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], dtype=np.uint32)
    key = np.array([[0, 0], [0xffffffff] * 2, [0xa4093822, 0x299f31d0]], dtype=np.uint32)
    rng = np.random.default_rng(0)
    ctr = np.concatenate([ctr, rng.integers(0, 2**32, (1000, 4), dtype=np.uint64).astype(np.uint32)])
    key = np.concatenate([key, rng.integers(0, 2**32, (1000, 2), dtype=np.uint64).astype(np.uint32)])
    dc = torch.as_tensor(ctr.view(np.int32), device="cuda")
    dk = torch.as_tensor(key.view(np.int32), device="cuda")
    out = torch.empty_like(dc)
    _lib.check(_lib.load().rmn_philox_raw(len(ctr), _lib.ptr(dc), _lib.ptr(dk), _lib.ptr(out), _lib.stream_ptr()))
    got = out.cpu().numpy().view(np.uint32)
    assert " ".join("%08x" % v for v in got[0]) == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert " ".join("%08x" % v for v in got[2]) == "d16cfe09 94fdcceb 5001e420 24126ea1"
    assert np.array_equal(got, philox.philox4x32_10(ctr, key))


def test_engine_draws_are_standard_normal_and_uniform():
    import torch
    from riemann_b200 import _lib
    from oracle import philox
    n, nn = 200000, 6
    z = torch.empty((n, nn), dtype=torch.float64, device="cuda")
    u = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().rmn_rng_draws(99, 1000, 5, n, nn, _lib.ptr(z), _lib.ptr(u), _lib.stream_ptr()))
    z, u = z.cpu().numpy(), u.cpu().numpy()
    assert abs(z.mean()) < 0.005 and abs(z.var() - 1) < 0.01
    assert abs(np.mean(z ** 4) - 3.0) < 0.05                        # kurtosis
    assert np.max(np.abs(np.corrcoef(z.T) - np.eye(nn))) < 0.01
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.003
    # the accept uniform is word 0 of block 0xFFFFFFFF of the numpy Philox
    bits = philox.draw_block(99, 1000 + np.arange(16), 5, 0xFFFFFFFF)[:, 0]
    assert np.array_equal(u[:16], philox.u01(bits))
