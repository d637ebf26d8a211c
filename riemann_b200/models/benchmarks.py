"""
Benchmark distributions -- same names and parameters as riemann/models/benchmarks.py:12-26.
Built lazily (each needs a CUDA device): ``benchmarks.benchmark_gauss2d_corr`` etc.
"""
import numpy as np

from .gaussian import MultiGaussianDist

_SPECS = {
    "benchmark_gauss1d": lambda: MultiGaussianDist(0.0, 1.0),
    "benchmark_gauss2d_iso": lambda: MultiGaussianDist(np.zeros(2), np.eye(2)),
    "benchmark_gauss2d_corr": lambda: MultiGaussianDist(
        np.zeros(2), 0.1 * np.eye(2) + 0.9 * np.ones((2, 2))),
    "benchmark_gauss100d_iso": lambda: MultiGaussianDist(np.zeros(100), np.eye(100)),
    "benchmark_gauss100d_corr": lambda: MultiGaussianDist(
        np.zeros(100), 0.1 * np.eye(100) + 0.9 * np.ones((100, 100))),
}
_cache = {}


def gauss_corr(d):
    """The 0.1 I + 0.9 11^T family at any d (config 3 uses d = 1000)."""
    return MultiGaussianDist(np.zeros(d), 0.1 * np.eye(d) + 0.9 * np.ones((d, d)))


def __getattr__(name):
    if name in _SPECS:
        if name not in _cache:
            _cache[name] = _SPECS[name]()
        return _cache[name]
    raise AttributeError(name)
