"""
CPU: analytic pins for the parts of the oracle the reference cannot pin
(logistic regression, simplified mMALA, Philox, ESS) -- SURVEY.md section 8c.
"""
import numpy as np

from oracle import riemann_port as port
from oracle import philox, ess


def test_philox_random123_kat():
    kat = [([0, 0, 0, 0], [0, 0], "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ([0xffffffff] * 4, [0xffffffff] * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, want in kat:
        out = philox.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32))
        assert " ".join("%08x" % v for v in out) == want


def test_philox_uniformity():
    bits = philox.draw_block(12345, np.arange(20000), 7, 0)
    u = philox.u01(bits).ravel()
    assert 0.0 < u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005


def _logistic(N=300, d=5):
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=5)
    return port.LogisticRegression(X, y, pv), ts


def test_logistic_gradient_finite_difference():
    m, ts = _logistic()
    th = ts * 0.7
    g = m.grad_log_posterior(th)
    h = 1e-6
    for j in range(len(th)):
        e = np.zeros_like(th)
        e[j] = h
        fd = (m.log_posterior(th + e) - m.log_posterior(th - e)) / (2 * h)
        assert abs(fd - g[j]) < 1e-5 * max(1.0, abs(g[j]))


def test_logistic_against_scikit_learn():
    """An independent third-party statement of the same model (the reference has none): scikit-learn's log_loss is
    -log-likelihood, and its L2-penalised LogisticRegression without intercept minimises 0.5 w.w + C sum(log-loss),
    i.e. its optimum is the posterior mode for prior_var = C -- where the oracle's gradient must vanish and its
    log-posterior must not be improved by the oracle's own Newton step."""
    import sklearn.linear_model
    import sklearn.metrics
    from scipy.special import expit
    X, y, ts, _ = port.make_logistic_problem(400, 6, seed=7)
    pv = 3.0
    m = port.LogisticRegression(X, y, pv)
    for th in (0.3 * ts, -1.2 * ts, np.zeros_like(ts)):
        ll = -sklearn.metrics.log_loss(y, expit(X @ th), normalize=False, labels=[0, 1])
        assert abs(m.log_likelihood(th) - ll) < 1e-9 * max(1.0, abs(ll))
    fit = sklearn.linear_model.LogisticRegression(C=pv, fit_intercept=False, tol=1e-12, max_iter=2000).fit(X, y)
    mode = fit.coef_.ravel()
    g = m.grad_log_posterior(mode)
    assert np.max(np.abs(g)) < 1e-5 * np.max(np.abs(m.grad_log_posterior(np.zeros_like(mode))))
    newton = mode + np.linalg.solve(m.metric(mode), g)
    assert np.max(np.abs(newton - mode)) < 1e-6
    assert m.log_posterior(mode) >= max(m.log_posterior(mode + 1e-3 * e) for e in np.eye(len(mode)))


def test_logistic_metric_spd_and_is_neg_hessian():
    m, ts = _logistic()
    th = ts * 0.3
    G = m.metric(th)
    assert np.allclose(G, G.T) and np.all(np.linalg.eigvalsh(G) > 0)
    h = 1e-5
    H = np.array([(m.grad_log_posterior(th + h * e) - m.grad_log_posterior(th - h * e)) / (2 * h)
                  for e in np.eye(len(th))])
    assert np.allclose(-H, G, rtol=1e-5, atol=1e-6)   # logistic: Fisher == -Hessian


def test_mmala_logq_antisymmetry_and_forward_term():
    m, ts = _logistic(N=200, d=4)
    prop = port.SimplifiedMMALA(0.8, m)
    a = ts * 0.5
    xi = np.array([0.3, -1.1, 0.4, 0.9])
    prop.draws = port.VectorTapeDraws(xi[None, :], np.zeros(1))
    b, lqr_ab = prop.propose(a)
    La, mean_a, ld_a = prop._geometry(a)
    # the forward residual is eps*xi by construction
    assert np.allclose(La.T @ (b - mean_a), prop.eps * xi, atol=1e-12)
    # reverse move b -> a uses the noise that maps back onto a
    Lb, mean_b, ld_b = prop._geometry(b)
    xi_rev = Lb.T @ (a - mean_b) / prop.eps
    prop.draws = port.VectorTapeDraws(xi_rev[None, :], np.zeros(1))
    a2, lqr_ba = prop.propose(b)
    assert np.allclose(a2, a, atol=1e-10)
    assert abs(lqr_ab + lqr_ba) < 1e-9


def test_mala_equals_textbook():
    """VanillaHMC(eps,1) == MALA with h = eps^2."""
    m = port.benchmark_gauss(3)
    eps = 0.3
    th = np.array([0.2, -0.4, 1.0])
    xi = np.array([0.5, -0.2, 1.3])
    prop = port.MALA(eps, m.grad_log_likelihood)
    prop.draws = port.VectorTapeDraws(xi[None, :], np.zeros(1))
    thp, lqr = prop.propose(th)
    g, gp = m.grad_log_likelihood(th), None
    mean = th + 0.5 * eps ** 2 * g
    assert np.allclose(thp, mean + eps * xi, atol=1e-14)
    gp = m.grad_log_likelihood(thp)
    lq_f = -np.sum((thp - mean) ** 2) / (2 * eps ** 2)
    lq_r = -np.sum((th - thp - 0.5 * eps ** 2 * gp) ** 2) / (2 * eps ** 2)
    assert abs(lqr - (lq_f - lq_r)) < 1e-12


def test_sampler_quirks():
    """Appendix A 1-3: strict <, min(0, nan) == 0 accepts -inf -> -inf, inf/nan -> -inf."""
    class Flat(port.Model):
        def log_prior(self, th):
            return 0.0

        def log_likelihood(self, th):
            return -np.inf if th[0] < 0 else (np.inf if th[0] > 10 else 0.0)

    class Shift(port.Proposal):
        def __init__(self, d):
            self.d = d

        def propose(self, th):
            return th + self.d, 0.0
    m = Flat()
    assert m.log_posterior(np.array([11.0])) == -np.inf          # +inf maps to -inf
    s = port.Sampler(m, Shift(-1.0), np.array([-5.0]),
                     draws=port.VectorTapeDraws(np.zeros((3, 1)), np.array([0.999999, 0.5, 0.5])))
    s.sample()
    assert s._chain_thetas[-1][0] == -6.0                         # -inf -> -inf accepted
    s2 = port.Sampler(m, Shift(-1.0), np.array([0.5]),
                      draws=port.VectorTapeDraws(np.zeros((1, 1)), np.array([1e-300])))
    s2.sample()
    assert s2._chain_thetas[-1][0] == 0.5                         # finite -> -inf rejected


def test_run_slicing_and_resume():
    """sampler.py:49-54: history = start + Nsamples, then [Nburn::Nthin]; run() resumes."""
    np.random.seed(3)
    s = port.Sampler(port.benchmark_gauss(2), port.MetropolisRandomWalk(0.5 * np.eye(2)), np.ones(2))
    s.run(100, 10, 3)
    assert len(s._chain_thetas) == len(range(10, 101, 3))
    last = s._chain_thetas[-1]
    s.run(5)
    assert np.array_equal(s._chain_thetas[0], last) and len(s._chain_thetas) == 6


def test_ess_estimators_agree_on_ar1():
    rng = np.random.default_rng(0)
    phi, n, K = 0.8, 4000, 256
    x = np.zeros((n, K))
    x[0] = rng.standard_normal(K) / np.sqrt(1 - phi ** 2)
    e = rng.standard_normal((n, K))
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    tau_true = (1 + phi) / (1 - phi)
    tau_sokal = ess.integrated_time_multi(x[:, :32, None])[0]
    ess_tot, tau_mom, rhat = ess.ess_from_chain_moments(n, x.mean(0)[:, None], x.var(0, ddof=1)[:, None])
    assert abs(tau_sokal - tau_true) / tau_true < 0.15
    assert abs(tau_mom[0] - tau_true) / tau_true < 0.25
    assert abs(rhat[0] - 1) < 0.01
