"""
Distributional gate, part 2 (BASELINE north_star: "long runs agree distributionally with the
reference sampler: posterior moments, KS on marginals").  Kolmogorov-Smirnov tests on the marginals
of many independent device chains (Philox mode) against

* the analytic target for the Gaussian benchmarks (riemann/models/benchmarks.py:18-26),
* thinned samples of 96 chains of the UNMODIFIED reference sampler on the changepoint problem
  (tests/golden/changepoint_marginals.npz, written by oracle/gen_golden.py; every second stored
  sample is used, lag 800 steps, where the autocorrelation is < 0.1),
* a long thinned CPU chain of the numpy oracle for the logistic model (no reference implementation).

Device samples are the states of K independent chains at one instant, so they are i.i.d. draws of
the chain's time-t law.  Runs are seeded: the outcome is deterministic.  Gate: p > 1e-3 per marginal
(Bonferroni-style: at most 8 marginals per test).
"""
import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu
P_MIN = 1e-3


def test_gauss2d_rw_marginals_are_standard_normal():
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    m = benchmarks.benchmark_gauss2d_corr
    K = 8192
    s = Sampler(m, MetropolisRandomWalk(0.5 * np.eye(2)), np.ones(2), K=K, seed=11)
    s.run(3000, trace=False)
    th = np.asarray(s._chain_thetas[-1])
    for j in range(2):
        assert stats.kstest(th[:, j], "norm").pvalue > P_MIN
    # the two principal axes of C = [[1, .9], [.9, 1]]: variances 1.9 and 0.1
    assert stats.kstest((th[:, 0] + th[:, 1]) / np.sqrt(2 * 1.9), "norm").pvalue > P_MIN
    assert stats.kstest((th[:, 0] - th[:, 1]) / np.sqrt(2 * 0.1), "norm").pvalue > P_MIN


def test_gauss100d_mala_marginals_are_standard_normal():
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.hamiltonian import MALA
    m = benchmarks.benchmark_gauss100d_corr
    K = 4096
    rng = np.random.default_rng(0)
    th0 = rng.standard_normal((K, 100)) * np.sqrt(0.1) + rng.standard_normal((K, 1)) * np.sqrt(0.9)
    s = Sampler(m, MALA(0.12, m.grad_log_likelihood), th0, seed=3)
    s.run(800, trace=False)
    th = np.asarray(s._chain_thetas[-1])
    for j in (0, 1, 50, 99):
        assert stats.kstest(th[:, j], "norm").pvalue > P_MIN
    # common mode: mean(theta) ~ N(0, 0.9 + 0.1/100); a stiff direction: (theta_0 - theta_1)/sqrt(0.2)
    assert stats.kstest(th.mean(1) / np.sqrt(0.9 + 0.1 / 100), "norm").pvalue > P_MIN
    assert stats.kstest((th[:, 0] - th[:, 1]) / np.sqrt(0.2), "norm").pvalue > P_MIN


def test_changepoint_marginals_match_reference_sampler(golden):
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams, ChangepointRegression1D
    from riemann_b200.proposals.changepoint import ChangepointRegression1DProp
    g = golden("changepoint_marginals")
    pm, pp, _, _ = port.make_changepoint_problem()
    dm = ChangepointRegression1D(pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb, int(pm.kmax), pm.alpha, pm.beta)
    dp = ChangepointRegression1DProp(dm, pp.hscale)
    K = 4096
    s = Sampler(dm, dp, ChangepointParams([2.0], [1.0, 3.0], 0.1), K=K, seed=77)
    s.run(int(g["T"]), trace=False)
    tr = s._chain_thetas
    k, cpx, cpv, sig = tr.k[-1], tr.cpx[-1], tr.cpv[-1], tr.sig[-1]
    ref_sig = g["sig"][:, ::2].ravel()
    assert stats.ks_2samp(sig, ref_sig).pvalue > P_MIN
    for qi, q in enumerate(g["query"]):
        xq = pm.xmin + (pm.xmax - pm.xmin) * (q + 0.5) / 6.0
        yq = np.array([cpv[c, np.searchsorted(cpx[c, :k[c]], xq)] for c in range(K)])
        assert stats.ks_2samp(yq, g["yq"][:, ::2, qi].ravel()).pvalue > P_MIN
    # k is discrete: chi-square of the device histogram against the reference frequencies
    ref_k = g["k"][:, ::2].ravel()
    lo, hi = 4, 10
    bins = np.arange(lo, hi + 2)
    dev_h = np.histogram(np.clip(k, lo, hi), bins)[0].astype(float)
    ref_h = np.histogram(np.clip(ref_k, lo, hi), bins)[0].astype(float)
    assert stats.chi2_contingency(np.array([dev_h, ref_h]))[1] > P_MIN


@pytest.mark.parametrize("kind", ["mala", "mmala"])
def test_logistic_marginals_match_oracle_chain(kind):
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.models.logistic import LogisticRegression
    from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA
    N, d = 400, 5
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=21)
    dm, om = LogisticRegression(X, y, pv), port.LogisticRegression(X, y, pv)
    eps = 0.3 if kind == "mala" else 0.9
    np.random.seed(5)
    op = port.MALA(eps, om.grad_log_posterior) if kind == "mala" else port.SimplifiedMMALA(eps, om)
    o = port.Sampler(om, op, ts.copy())
    o.run(13000, 1000, 20)                              # 600 samples, lag 20 steps
    och = np.array(o._chain_thetas)
    x = och - och.mean(0)
    assert np.all(np.abs((x[1:] * x[:-1]).mean(0) / x.var(0)) < 0.2)     # thinned samples ~ independent
    p = MALA(eps, dm.grad_log_posterior) if kind == "mala" else SimplifiedMMALA(eps, dm)
    s = Sampler(dm, p, ts.copy(), K=2048, seed=17)
    s.run(600, trace=False)
    th = np.asarray(s._chain_thetas[-1])
    for j in range(d):
        assert stats.ks_2samp(th[:, j], och[:, j]).pvalue > P_MIN
