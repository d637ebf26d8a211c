"""
TEST INFRASTRUCTURE -- numpy Philox4x32-10 (Salmon et al., SC'11; Random123).

Restates the published algorithm; pinned in tests/test_oracle_philox.py against
the Random123 known-answer vectors.  It is the checker for the device RNG in
riemann_b200/csrc/philox.cuh, whose counter/key convention it mirrors:

    key     = (seed_lo, seed_hi)
    counter = (block, step_lo, step_hi, global_chain_id)

The reference itself uses numpy's global MT19937 stream (randomwalk.py:25,
sampler.py:84); counter-based Philox is the engine's replacement.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """
    counter: uint32 array [..., 4]; key: uint32 array [..., 2] (broadcastable).
    Returns uint32 [..., 4].
    """
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(key, dtype=np.uint32, copy=True)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0 = np.broadcast_to(k[..., 0], c0.shape).copy()
    k1 = np.broadcast_to(k[..., 1], c0.shape).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1)


def draw_block(seed, chain, step, block):
    """uint32[..., 4] for the engine's (seed, chain, step, block) convention."""
    chain = np.asarray(chain, dtype=np.uint64)
    step = np.asarray(step, dtype=np.uint64)
    block = np.asarray(block, dtype=np.uint64)
    shape = np.broadcast(chain, step, block).shape
    ctr = np.empty(shape + (4,), dtype=np.uint32)
    ctr[..., 0] = np.broadcast_to(block & MASK32, shape)
    ctr[..., 1] = np.broadcast_to(step & MASK32, shape)
    ctr[..., 2] = np.broadcast_to(step >> np.uint64(32), shape)
    ctr[..., 3] = np.broadcast_to(chain & MASK32, shape)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key)


def u01(bits):
    """The engine's uniform on (0,1): (x + 0.5) * 2^-32 in fp64."""
    return (np.asarray(bits, dtype=np.float64) + 0.5) * (1.0 / 4294967296.0)
