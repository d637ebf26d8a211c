"""
CPU, property-based (hypothesis, derandomised): analytic identities the reference's algorithm satisfies, checked on
the oracle restatement.  The reference has no assertions of its own, so besides the recorded
fixtures these identities are what pins the port: reversibility of the trans-dimensional maps and of the leapfrog,
the pCN prior-reversibility identity, the whitened kinetic energy, and the prefix-sum form of the changepoint
likelihood that the device kernel evaluates.
"""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import riemann_port as port

SET = dict(max_examples=60, deadline=None, derandomize=True)


def _theta(rng, k, xmin=1.0, xmax=3.0):
    cpx = np.sort(rng.uniform(xmin, xmax, k))
    cpv = rng.uniform(0.5, 3.0, k + 1)
    return port.ChangepointParams(cpx, cpv, 0.1 + rng.uniform())


def _model(rng, M=60):
    x = np.sort(rng.uniform(1.0, 3.0, M))
    y = rng.normal(size=M) + 2.0
    return port.ChangepointRegression1D(x, y, 1.0, 3.0, 5.0, 10, 1, 1)


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), k=st.integers(0, 7))
def test_birth_then_death_is_the_identity_and_the_jacobians_cancel(seed, k):
    """changepoint.py:193-240: subtract_changepoint undoes add_changepoint; log|J| + log|J^-1| = 0."""
    rng = np.random.default_rng(seed)
    m, th = _model(rng), _theta(rng, k)
    s, u = rng.uniform(1.0, 3.0), rng.uniform(0.05, 0.95)
    th2, lj = m.add_changepoint(th, s, u)
    n = int(np.searchsorted(th.cpx, s))
    assert len(th2.cpx) == k + 1 and th2.cpx[n] == s and np.all(np.diff(th2.cpx) >= 0)
    th3, lj_back = m.subtract_changepoint(th2, n)
    assert np.array_equal(th3.cpx, th.cpx)
    assert np.allclose(th3.cpv, th.cpv, rtol=1e-13, atol=0)
    assert abs(lj + lj_back) < 1e-12 * max(1.0, abs(lj))


@settings(**SET)
@given(h=st.floats(0.2, 5.0), u=st.floats(0.05, 0.95))
def test_birth_jacobian_matches_finite_differences(h, u):
    """The analytic log|det d(h1,h2)/d(h,u)| = log(h / (u (1 - u))) that replaces autograd.jacobian (changepoint.py:72-78)."""
    def fmap(h_, u_):
        f = np.sqrt((1 - u_) / u_)
        return np.array([h_ / f, h_ * f])
    e = 1e-6
    J = np.column_stack([(fmap(h + e, u) - fmap(h - e, u)) / (2 * e), (fmap(h, u + e) - fmap(h, u - e)) / (2 * e)])
    assert abs(np.log(abs(np.linalg.det(J))) - port.birth_map_logjac(h, u)) < 1e-6


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), k=st.integers(0, 9))
def test_prefix_sum_likelihood_equals_the_direct_sum(seed, k):
    """The device kernel's O(k log M) form: per run of data between changepoints, n v^2 - 2 v S1 + S2 from prefix sums
    of the centred responses (riemann_b200/csrc/changepoint.cu) == the reference's O(M) residual sum (changepoint.py:106-126)."""
    rng = np.random.default_rng(seed)
    m, th = _model(rng, M=int(rng.integers(5, 130))), _theta(rng, k)
    yc = m.y.mean()
    cy = np.concatenate([[0.0], np.cumsum(m.y - yc)])
    cyy = np.concatenate([[0.0], np.cumsum((m.y - yc) ** 2)])
    bounds = np.concatenate([[0], np.searchsorted(m.x, th.cpx, side="right"), [len(m.x)]])
    ss = 0.0
    for j in range(k + 1):
        lo, hi = bounds[j], bounds[j + 1]
        v = th.cpv[j] - yc
        ss += (hi - lo) * v * v - 2.0 * v * (cy[hi] - cy[lo]) + (cyy[hi] - cyy[lo])
    M, s2 = len(m.x), th.sig ** 2
    logl = -0.5 * (ss / s2 + M * np.log(s2) + M * np.log(2 * np.pi))
    ref = m.log_likelihood(th)
    assert abs(logl - ref) < 1e-9 * max(1.0, abs(ref))


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), d=st.integers(1, 6), rho=st.floats(0.05, 0.98))
def test_pcn_is_reversible_with_respect_to_its_gaussian(seed, d, rho):
    """randomwalk.py:88-100: pi0(theta) q(theta'|theta) = pi0(theta') q(theta|theta') for pi0 = N(0, C), so the
    returned logqratio = log q(theta'|theta) - log q(theta|theta') equals log pi0(theta') - log pi0(theta)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    C = A @ A.T / d + 0.3 * np.eye(d)
    p = port.pCN(C, rho)
    p.draws = port.LiveDraws()
    np.random.seed(seed % 2 ** 31)
    th = rng.standard_normal(d)
    thp, lqr = p.propose(th)
    prior = port.MultiGaussianDist(np.zeros(d), C)
    want = prior.log_likelihood(thp) - prior.log_likelihood(th)
    assert abs(lqr - want) < 1e-9 * max(1.0, abs(want))


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), d=st.integers(1, 6), nsteps=st.integers(1, 6), mass=st.booleans())
def test_leapfrog_is_reversible_and_hmc_returns_the_kinetic_energy_difference(seed, d, nsteps, mass):
    """hamiltonian.py:13-52, 76-91: flipping the final momentum and integrating again returns to the start; the
    proposal's logqratio is K(p') - K(p0) with K(p) = p^T M^-1 p / 2 (computed there through chM^-1 p)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    C = A @ A.T / d + 0.3 * np.eye(d)
    g = port.MultiGaussianDist(rng.standard_normal(d), C)
    M = None
    if mass:
        B = rng.standard_normal((d, d))
        M = B @ B.T / d + 0.5 * np.eye(d)
    q0, p0 = rng.standard_normal(d), rng.standard_normal(d)
    eps = 0.1
    p1, q1 = port.leapfrog(p0, q0, nsteps, eps, g.grad_log_likelihood, M)
    pb, qb = port.leapfrog(-p1, q1, nsteps, eps, g.grad_log_likelihood, M)
    assert np.allclose(qb, q0, rtol=0, atol=1e-10) and np.allclose(-pb, p0, rtol=0, atol=1e-10)

    prop = port.VanillaHMC(eps, nsteps, g.grad_log_likelihood, M=M)
    xi = rng.standard_normal(d)
    prop.draws = port.VectorTapeDraws(xi[None, :], np.array([0.5]))
    thp, lqr = prop.propose(q0)
    pstart = xi if M is None else np.linalg.cholesky(M) @ xi
    pend, qend = port.leapfrog(pstart, q0, nsteps, eps, g.grad_log_likelihood, M)
    Minv = np.eye(d) if M is None else np.linalg.inv(M)
    want = 0.5 * (pend @ Minv @ pend - pstart @ Minv @ pstart)
    assert np.allclose(thp, qend, rtol=0, atol=1e-12)
    assert abs(lqr - want) < 1e-9 * max(1.0, abs(want))
