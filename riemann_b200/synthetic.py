"""
Synthetic inputs of the measurement plan, plain numpy: the data sets bench.py, smoke() and
the scripts feed to the device engine.  The oracle's `make_*_problem` helpers wrap the same arrays in the numpy
port's classes, so the CPU baseline and the engine see identical inputs.  numpy's Philox bit generator makes them
reproducible across numpy versions.  Nothing here runs on the sampling path.
"""
import numpy as np

SEED_BASE = 20261018


def changepoint_problem(seed=SEED_BASE + 2, Ncpx=5, Ndata=100, xmin=1.0, xmax=3.0, hmin=1.0, hmax=3.0, sig=0.1):
    """Recipe of examples/test_changepoint.py:137-150,167 with a seeded generator.  Returns a dict: data `x`, `y`;
    model arguments `xmin, xmax, lamb, kmax, alpha, beta` (changepoint.py:86-104); proposal `hscale`
    (test_changepoint.py:24-32); start state `theta0 = (cpx, cpv, sig)` and the generating `theta_true`."""
    rng = np.random.Generator(np.random.Philox(seed))
    cpx = np.sort(rng.uniform(xmin, xmax, size=Ncpx))
    cpv = rng.uniform(hmin, hmax, size=Ncpx + 1)
    x = np.sort(xmin + (xmax - xmin) * rng.uniform(size=Ndata))
    y = cpv[np.searchsorted(cpx, x)] + sig * rng.normal(size=Ndata)        # predict(), changepoint.py:175-181
    return dict(x=x, y=y, xmin=xmin, xmax=xmax, lamb=1.0 * Ncpx, kmax=2 * Ncpx, alpha=1, beta=1, hscale=hmax - hmin,
                theta0=(np.array([0.5 * (xmin + xmax)]), np.array([hmin, hmax]), 0.1),
                theta_true=(cpx, cpv, sig))


def logistic_problem(N, d, seed=SEED_BASE + 4, prior_var=100.0, dtype=np.float64):
    """Configs 4 / 5: X_ij ~ N(0, 1/d), theta* ~ N(0, I), y_i ~ Bernoulli(sigmoid(x_i . theta*)), prior N(0, prior_var I).
    Returns (X, y, theta_star, prior_var)."""
    from scipy.special import expit
    rng = np.random.Generator(np.random.Philox(seed))
    X = (rng.standard_normal((N, d)) / np.sqrt(d)).astype(dtype)
    theta_star = rng.standard_normal(d)
    y = (rng.uniform(size=N) < expit(X.astype(np.float64) @ theta_star)).astype(np.float64)
    return X, y, theta_star, prior_var
