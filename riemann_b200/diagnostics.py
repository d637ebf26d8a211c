"""
Host-side chain diagnostics for traced chains: integrated autocorrelation time with Sokal's
automatic window, the estimator `emcee.autocorr.integrated_time` implements and the reference calls
at examples/test_randomwalk.py:42 to report "steps per independent sample" (emcee itself is an
un-vendored, un-pinned dependency of the reference and is not installed here: the algorithm is
restated from its published description, Goodman & Weare 2010 / Sokal 1989, and checked against the
analytic tau of AR(1) chains in tests/test_abi_and_host.py).

    rho(t)  = autocorrelation function (FFT, zero-padded to 2 * next_pow2(N))
    tau(W)  = 1 + 2 sum_{t=1..W} rho(t)
    window  = smallest W with W >= c * tau(W)         (c = 5)

`integrated_time(x)` takes one chain [N] or [N, d]; `integrated_time_chains(x)` takes K chains
[N, K] or [N, K, d] and averages the autocorrelation FUNCTION over chains before windowing (what emcee
>= 3 does with walkers), which is the low-variance estimator when many short chains are available --
the engine's case.  The diagnostics block of the device engine (riemann_b200.distributed
.summarize_block) uses the moment-based many-chain estimator tau = n B / W instead, which needs no
trace; bench.py reports both for a traced subset.
"""
import numpy as np


def _acf(x):
    """Autocorrelation function along axis 0 of x [N, ...] (biased, normalised to rho(0) = 1)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    m = 1
    while m < n:
        m <<= 1
    f = np.fft.rfft(x - x.mean(axis=0, keepdims=True), n=2 * m, axis=0)
    acov = np.fft.irfft(f * np.conjugate(f), n=2 * m, axis=0)[:n]
    return acov


def _window(taus, c):
    ok = np.arange(len(taus)) < c * taus
    return int(np.argmin(ok)) if np.any(ok) and not np.all(ok) else len(taus) - 1


def _tau_from_acov(acov, c):
    rho = acov / acov[0]
    taus = 2.0 * np.cumsum(rho) - 1.0
    return float(taus[_window(taus, c)])


def integrated_time(x, c=5.0):
    """One chain x [N] or [N, d] -> tau per column (in steps of the chain)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    acov = _acf(x)
    return np.array([_tau_from_acov(acov[:, j], c) for j in range(x.shape[1])])


def integrated_time_device(x, c=5.0, return_window=False):
    """The same estimator on the GPU (`rmn_autocorr_tau`: batched cuFFT + windowing kernels).  x is a trace
    [N, K] or [N, K, d], a numpy array (uploaded) or a float64 CUDA tensor (used in place, e.g. the device trace
    of a run).  Returns tau per functional (and the windows)."""
    import ctypes as C
    from . import _lib
    torch = _lib.require_cuda()
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")
    if t.dtype != torch.float64 or not t.is_cuda:
        raise ValueError("integrated_time_device needs float64 data on the GPU")
    if t.ndim == 2:
        t = t[:, :, None]
    t = t.contiguous()
    n, K, nd = t.shape
    tau = np.empty(nd, dtype=np.float64)
    win = np.empty(nd, dtype=np.int64)
    _lib.check(_lib.load().rmn_autocorr_tau(_lib.ptr(t), n, K, nd, float(c), tau.ctypes.data_as(C.c_void_p),
                                            win.ctypes.data_as(C.c_void_p), _lib.stream_ptr()))
    return (tau, win) if return_window else tau


def integrated_time_chains(x, c=5.0, device=False):
    """K chains x [N, K] or [N, K, d] -> tau per functional, ACF averaged over the chains.
    device=True evaluates it on the GPU (integrated_time_device)."""
    if device:
        return integrated_time_device(x, c)
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    acov = _acf(x)                            # each chain centred on its own mean
    rho = (acov / acov[0:1]).mean(axis=1)     # normalised per chain, then averaged (emcee >= 3)
    return np.array([_tau_from_acov(rho[:, j], c) for j in range(x.shape[2])])


def effective_sample_size(x, c=5.0):
    """Pooled ESS of K traced chains [N, K(, d)]: K N / tau."""
    x = np.asarray(x)
    return x.shape[0] * x.shape[1] / integrated_time_chains(x, c)
