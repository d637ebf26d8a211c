"""
rmn_autocorr_tau: emcee's integrated autocorrelation time (examples/test_randomwalk.py:42) on the device trace
(SURVEY.md 8f N4), against its host twin riemann_b200/diagnostics.py -- which tests/test_abi_and_host.py pins to the
oracle restatement and to the analytic tau of AR(1) chains.  Both are fp64 FFT autocorrelations (cuFFT vs pocketfft):
tau within 1e-9, identical windows.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ar1(n, K, phis, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, K, len(phis)))
    e = rng.standard_normal((n, K, len(phis)))
    for j, phi in enumerate(phis):
        x[0, :, j] = e[0, :, j] / np.sqrt(1 - phi * phi)
        for t in range(1, n):
            x[t, :, j] = phi * x[t - 1, :, j] + e[t, :, j]
    return x


@pytest.mark.parametrize("n,K", [(5000, 16), (4096, 3), (777, 1)])
def test_device_tau_matches_host_and_ar1(n, K):
    from riemann_b200 import diagnostics as dg
    phis = (0.0, 0.5, 0.9)
    x = _ar1(n, K, phis, seed=n + K)
    host = dg.integrated_time_chains(x)
    dev, win = dg.integrated_time_device(x, return_window=True)
    assert np.allclose(dev, host, rtol=1e-9, atol=0), (dev, host)
    assert np.array_equal(dg.integrated_time_chains(x, device=True), dev)
    # the window is the first lag W >= 5 tau(W)
    assert np.all(win >= 5 * dev - 1e-9) and np.all(win < n)
    if n * K >= 10000:
        exact = np.array([(1 + p) / (1 - p) for p in phis])
        assert np.all(np.abs(dev / exact - 1) < 0.2), (dev, exact)
    if K == 1:
        assert np.allclose(dev, dg.integrated_time(x[:, 0, :]), rtol=1e-9, atol=0)


def test_layouts_other_c_and_degenerate_series():
    import torch
    from riemann_b200 import diagnostics as dg
    x = _ar1(1500, 5, (0.7,), seed=2)
    host = dg.integrated_time_chains(x[:, :, 0], c=3.0)
    assert np.allclose(dg.integrated_time_device(x[:, :, 0], c=3.0), host, rtol=1e-9)                    # [N, K]
    t = torch.as_tensor(x, device="cuda")
    assert np.allclose(dg.integrated_time_device(t, c=3.0), host, rtol=1e-9)                             # CUDA tensor in place
    # no window before the last lag: a strongly correlated short series
    y = np.cumsum(np.random.default_rng(0).standard_normal((40, 2, 1)), axis=0)
    dev, win = dg.integrated_time_device(y, return_window=True)
    assert np.allclose(dev, dg.integrated_time_chains(y), rtol=1e-9)
    # a constant chain has no autocorrelation function: nan on both sides
    z = np.ones((64, 2, 1))
    with np.errstate(all="ignore"):
        assert np.isnan(dg.integrated_time_chains(z)[0])
    assert np.isnan(dg.integrated_time_device(z)[0])
    with pytest.raises(Exception):
        dg.integrated_time_device(np.zeros((1, 2, 1)))


def test_tau_of_a_device_run_trace():
    """The use the reference makes of it: tau of a sampled chain (test_randomwalk.py:39-46), here 64 chains."""
    from riemann_b200 import Sampler, diagnostics as dg
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    s = Sampler(benchmarks.benchmark_gauss2d_corr, MetropolisRandomWalk(0.5 * np.eye(2)), np.zeros(2), K=64, seed=5)
    s.run(6000, 1000, 1)
    chain = np.asarray(s._chain_thetas)                     # [5001, 64, 2]
    assert chain.shape == (5001, 64, 2)
    host = dg.integrated_time_chains(chain)
    dev = dg.integrated_time_device(chain)
    assert np.allclose(dev, host, rtol=1e-9)
    assert np.all(dev > 2) and np.all(dev < 500)
