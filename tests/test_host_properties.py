"""CPU, property-based (hypothesis, derandomised): host-side logic of the drop-in layer that needs no GPU --
the canonical padded state layout of the changepoint model, chain / row sharding, and the summable diagnostics
block (combining the blocks of two ranks equals the block of their union)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

SET = dict(max_examples=80, deadline=None, derandomize=True)


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 12))
def test_pack_unpack_roundtrip_and_canonical_padding(seed, n):
    from riemann_b200.models.changepoint import LANES, ChangepointParams, pack_states, unpack_state
    rng = np.random.default_rng(seed)
    thetas = []
    for _ in range(n):
        k = int(rng.integers(0, LANES))
        thetas.append(ChangepointParams(np.sort(rng.uniform(1, 3, k)), rng.uniform(0.5, 3, k + 1), rng.uniform(0.05, 1)))
    k, cpx, cpv, sig = pack_states(thetas)
    assert cpx.shape == cpv.shape == (n, LANES) and k.dtype == np.int32
    for i, th in enumerate(thetas):
        assert np.all(cpx[i, k[i]:] == 0.0) and np.all(cpv[i, k[i] + 1:] == 0.0)      # zeros beyond the extent
        back = unpack_state(k[i], cpx[i], cpv[i], sig[i])
        assert np.array_equal(back.cpx, th.cpx) and np.array_equal(back.cpv, th.cpv) and back.sig == th.sig


def test_too_many_changepoints_and_bad_shapes_are_refused():
    from riemann_b200 import ParameterError
    from riemann_b200.models.changepoint import LANES, ChangepointParams, pack_states
    with pytest.raises(ParameterError):
        pack_states([ChangepointParams(np.linspace(1, 3, LANES), np.ones(LANES + 1), 0.1)])
    with pytest.raises(ValueError):                       # changepoint.py:35-39
        ChangepointParams([1.0, 2.0], [1.0, 2.0], 0.1)


@settings(**SET)
@given(total=st.integers(0, 10 ** 7), world=st.integers(1, 64))
def test_shards_partition_the_axis(total, world):
    from riemann_b200.distributed import shard_chains, shard_rows
    parts = [shard_chains(total, r, world) for r in range(world)]
    assert parts == [shard_rows(total, r, world) for r in range(world)]
    assert parts[0][0] == 0 and sum(k for _, k in parts) == total
    for (o0, k0), (o1, _) in zip(parts, parts[1:]):
        assert o1 == o0 + k0
    sizes = [k for _, k in parts]
    assert max(sizes) - min(sizes) <= 1                   # balanced to within one


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), ka=st.integers(2, 40), kb=st.integers(2, 40), n=st.integers(5, 60))
def test_diagnostics_blocks_are_summable_over_ranks(seed, ka, kb, n):
    """The all-reduced block of two ranks gives the statistics of the pooled chains (what makes one
    all_reduce(SUM) enough, distributed.reduce_block)."""
    from riemann_b200.distributed import summarize_block
    rng = np.random.default_rng(seed)
    nd = 3

    def block(x):                                          # x[n][K][nd] -> the layout summarize_block documents
        K = x.shape[1]
        m = x.mean(axis=0)                                 # chain means [K][nd]
        v = x.var(axis=0)                                  # biased within-chain variance
        return np.concatenate([[K, n, 0.3 * K * n, 0, n, 0], m.sum(0), (m * m).sum(0), v.sum(0)])

    xa = rng.standard_normal((n, ka, nd)) + rng.standard_normal((1, ka, nd))
    xb = rng.standard_normal((n, kb, nd)) + rng.standard_normal((1, kb, nd))
    ba, bb = block(xa), block(xb)
    summed = ba + bb
    summed[[1, 4]] = ba[[1, 4]]                            # samples / steps per chain are not additive
    pooled = summarize_block(block(np.concatenate([xa, xb], axis=1)))
    got = summarize_block(summed)
    for key in ("mean", "var", "tau", "ess", "rhat"):
        assert np.allclose(got[key], pooled[key], rtol=1e-10, atol=1e-12), key
    assert got["chains"] == ka + kb and abs(got["accept_rate"] - 0.3) < 1e-12
