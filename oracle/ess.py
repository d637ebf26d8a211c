"""
TEST INFRASTRUCTURE -- integrated autocorrelation time / ESS estimators.

``integrated_time`` restates the algorithm of emcee.autocorr.integrated_time
(Goodman & Weare 2010 / Sokal window; emcee is the un-vendored, un-pinned
dependency the reference calls at examples/test_randomwalk.py:42).  emcee is NOT
installed here, so this is "parity unpinned": no reference test fixes any tau.

``ess_from_chain_moments`` is the many-chain estimator the engine's diagnostics
block uses (variance of chain means vs. within-chain variance); the test-suite
checks the two agree on synthetic AR(1) chains.
"""
import numpy as np


def _next_pow_two(n):
    i = 1
    while i < n:
        i <<= 1
    return i


def autocorr_func_1d(x):
    x = np.asarray(x, dtype=np.float64)
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    return acf / acf[0]


def _auto_window(taus, c):
    m = np.arange(len(taus)) < c * taus
    return int(np.argmin(m)) if np.any(m) else len(taus) - 1


def integrated_time(chain, c=5.0):
    """
    chain: (N,) or (N, d) -- one chain.  Returns tau per column:
    tau(W) = 2 sum_{t<=W} rho(t) - 1 at the smallest W with W >= c tau(W).
    """
    chain = np.asarray(chain, dtype=np.float64)
    if chain.ndim == 1:
        chain = chain[:, None]
    out = np.empty(chain.shape[1])
    for j in range(chain.shape[1]):
        rho = autocorr_func_1d(chain[:, j])
        taus = 2.0 * np.cumsum(rho) - 1.0
        out[j] = taus[_auto_window(taus, c)]
    return out


def integrated_time_multi(chains, c=5.0):
    """chains: (N, W, d): average the ACF over W chains first (emcee >= 3 style)."""
    chains = np.asarray(chains, dtype=np.float64)
    N, W, d = chains.shape
    out = np.empty(d)
    for j in range(d):
        rho = np.zeros(N)
        for w in range(W):
            rho += autocorr_func_1d(chains[:, w, j])
        rho /= W
        taus = 2.0 * np.cumsum(rho) - 1.0
        out[j] = taus[_auto_window(taus, c)]
    return out


def ess_from_chain_moments(n, chain_mean, chain_var):
    """
    K independent chains of n post-burn-in steps each; per-chain means/variances of
    shape (K, d).  Var(chain mean) ~ sigma^2 tau / n  =>  tau = n B / W with
    B = variance of the chain means, W = mean within-chain variance.
    Returns (ess_total[d], tau[d], rhat[d]).
    """
    chain_mean = np.asarray(chain_mean, dtype=np.float64)
    chain_var = np.asarray(chain_var, dtype=np.float64)
    K = chain_mean.shape[0]
    B = np.var(chain_mean, axis=0, ddof=1)
    W = np.mean(chain_var, axis=0)
    tau = np.maximum(n * B / W, 1e-300)
    var_plus = (n - 1.0) / n * W + B
    rhat = np.sqrt(var_plus / W)
    return K * n / tau, tau, rhat
