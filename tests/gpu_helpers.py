"""Shared helpers for the -m gpu parity tests (device vs oracle / golden fixtures)."""
import numpy as np


def relerr(a, b):
    """max |a-b| / max(1, |b|) over finite entries; inf/nan patterns must agree."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    fa, fb = np.isfinite(a), np.isfinite(b)
    assert np.array_equal(fa, fb), "finite/non-finite pattern differs"
    assert np.array_equal(a[~fa], b[~fb]) or np.all(np.isnan(a[~fa]) == np.isnan(b[~fb]))
    if not fa.any():
        return 0.0
    return float(np.max(np.abs(a[fa] - b[fb]) / np.maximum(1.0, np.abs(b[fb]))))


def device_gauss(g, d):
    from riemann_b200.models.gaussian import MultiGaussianDist
    if "C" in g:
        return MultiGaussianDist(g["mu"], g["C"])
    C = 0.1 * np.eye(d) + 0.9 * np.ones((d, d)) if d > 1 else np.eye(1)
    return MultiGaussianDist(np.zeros(d), C)


def oracle_gauss(g, d):
    from oracle import riemann_port as port
    if "C" in g:
        return port.MultiGaussianDist(g["mu"], g["C"])
    return port.benchmark_gauss(d, corr=(d > 1))
