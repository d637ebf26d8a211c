"""
TEST INFRASTRUCTURE -- numpy fp64 restatement of riemann's MH hot path.

Never imported by the product (``riemann_b200``).  Used by tests/, smoke() and
bench.py's CPU-baseline legs as the checker / reported baseline.

Every class states the reference file:line whose arithmetic it restates.  The
restatement keeps the reference's *draw order* from numpy's global legacy
stream, so under ``np.random.seed(s)`` it reproduces the reference's chains
(checked in tests/test_oracle_vs_golden.py against fixtures generated from the
reference itself by oracle/gen_golden.py).

All randomness goes through a small "draw source" so the same code can
(a) consume the live numpy stream exactly like the reference, or
(b) replay a recorded tape -- the same tape the CUDA engine replays in its
    injected mode (layout constants below are shared with include/riemann_b200.h).
"""
import math

import numpy as np
from scipy.special import gammaln, expit
from scipy.linalg.lapack import dtrtrs


def solve_triangular(a, b, lower, check_finite=False):
    """scipy.linalg.solve_triangular without its per-call argument validation (11 us of Python per call, more than the
    whole d = 2 solve): the same LAPACK dtrtrs call scipy makes, transposed system for a C-ordered factor exactly like
    scipy's `_solve_triangular`, so results are bit-identical.  Non-finite inputs give non-finite outputs, as the
    reference's np.linalg.solve does (gaussian.py:51), instead of scipy's ValueError."""
    if a.flags.f_contiguous:
        x, info = dtrtrs(a, b, lower=lower, trans=0)
    else:
        x, info = dtrtrs(a.T, b, lower=not lower, trans=1)
    if info != 0:
        raise np.linalg.LinAlgError("dtrtrs: singular triangular factor (info = %d)" % info)
    return x

# --------------------------------------------------------------------------
# tape layout for the changepoint proposal (mirrors RMN_CP_SLOT_* in
# include/riemann_b200.h).  One row of NSLOT doubles per MH step.
# --------------------------------------------------------------------------
CP_SLOT_SEL1 = 0      # first block-selection uniform        test_changepoint.py:48
CP_SLOT_SEL2 = 1      # second (drawn only if first >= .20)  :51
CP_SLOT_SEL3 = 2      # third (drawn only if second >= .40)  :54
CP_SLOT_BD = 3        # birth/death coin (only if k>0)       :59
CP_SLOT_S = 4         # new changepoint location U(xmin,xmax) :60
CP_SLOT_DU = 5        # U(-0.1, 0.1)                          :61
CP_SLOT_N = 6         # randint(k) as a double                :67
CP_SLOT_ACC = 7       # accept uniform                        sampler.py:84
CP_SLOT_XI = 8        # normals xi_0 .. xi_{LANES-1}


def cp_nslot(lanes=16):
    return CP_SLOT_XI + lanes


class ParameterError(Exception):
    """Stands in for riemann.sampling_errors.ParameterError (:24-28)."""


# --------------------------------------------------------------------------
# draw sources
# --------------------------------------------------------------------------
class LiveDraws(object):
    """numpy's global legacy stream, consumed in the reference's order."""

    def normal(self, role, n):
        return np.random.normal(size=(n,))

    def uniform(self, role, low=0.0, high=1.0):
        if low == 0.0 and high == 1.0:
            return np.random.uniform()
        return np.random.uniform(low, high)

    def randint(self, role, n):
        return np.random.randint(n)

    def next_step(self):
        pass


class VectorTapeDraws(object):
    """Replay for fixed-d proposals: xi[T, d] normals and u[T] accept uniforms."""

    def __init__(self, xi, u):
        self.xi = np.asarray(xi, dtype=np.float64)
        self.u = np.asarray(u, dtype=np.float64)
        self.t = 0

    def normal(self, role, n):
        return self.xi[self.t, :n].copy()

    def uniform(self, role, low=0.0, high=1.0):
        return self.u[self.t]

    def randint(self, role, n):
        raise NotImplementedError

    def next_step(self):
        self.t += 1


class SlotTapeDraws(object):
    """Replay for the changepoint proposal: tape[T, NSLOT] (see CP_SLOT_*)."""
    _slot = {"sel1": CP_SLOT_SEL1, "sel2": CP_SLOT_SEL2, "sel3": CP_SLOT_SEL3,
             "bd": CP_SLOT_BD, "s": CP_SLOT_S, "du": CP_SLOT_DU, "acc": CP_SLOT_ACC}

    def __init__(self, tape):
        self.tape = np.asarray(tape, dtype=np.float64)
        self.t = 0

    def normal(self, role, n):
        return self.tape[self.t, CP_SLOT_XI:CP_SLOT_XI + n].copy()

    def uniform(self, role, low=0.0, high=1.0):
        return self.tape[self.t, self._slot[role]]       # stored already scaled

    def randint(self, role, n):
        return int(self.tape[self.t, CP_SLOT_N])

    def next_step(self):
        self.t += 1


# --------------------------------------------------------------------------
# Model protocol                                   riemann/models/model.py:4-64
# --------------------------------------------------------------------------
class Model(object):
    def log_likelihood(self, theta):
        raise NotImplementedError("abstract")

    def log_prior(self, theta):
        raise NotImplementedError("abstract")

    def log_posterior(self, theta):
        # model.py:49-54 -- ANY inf (either sign) or nan in either term => -inf
        lp = self.log_prior(theta)
        ll = self.log_likelihood(theta)
        bad = not (math.isfinite(float(np.squeeze(lp))) and
                   math.isfinite(float(np.squeeze(ll))))
        return -np.inf if bad else lp + ll

    def logL(self, theta):
        return self.log_likelihood(theta)

    def logP(self, theta):
        return self.log_prior(theta)

    def __call__(self, theta):
        return self.log_posterior(theta)


class MultiGaussianDist(Model):
    """Dense Gaussian with known mean/cov.       riemann/models/gaussian.py:21-58"""

    def __init__(self, mu, C):
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        C = np.atleast_2d(np.asarray(C, dtype=np.float64))
        if C.shape[0] != C.shape[1]:                         # gaussian.py:35-36
            raise ParameterError("C has non-square shape {}".format(C.shape))
        if C.shape[1] != mu.shape[0]:                        # gaussian.py:37-39
            raise ParameterError("mu and C have incompatible shapes")
        self.mu = mu
        self.C = C
        self.L = np.linalg.cholesky(C)                       # gaussian.py:42
        self.logdetC = 2 * np.sum(np.log(np.diag(self.L)))   # gaussian.py:43
        self.Ndim = len(mu)

    def log_prior(self, theta):
        return 0.0                                           # gaussian.py:46-47

    def log_likelihood(self, theta):
        # gaussian.py:49-52; the reference uses a general solve on the triangular
        # factor, a triangular solve gives the same u to fp64 round-off
        y = theta - self.mu
        u = solve_triangular(self.L, y, lower=True, check_finite=False)
        return -0.5 * (np.dot(u, u) + len(y) * np.log(2 * np.pi) + self.logdetC)

    def grad_log_likelihood(self, theta):
        # gaussian.py:54-58:  -C^{-1} (theta - mu)
        y = theta - self.mu
        u = solve_triangular(self.L, y, lower=True, check_finite=False)
        return -solve_triangular(self.L.T, u, lower=False, check_finite=False)

    grad_log_posterior = grad_log_likelihood                 # prior is flat


def benchmark_gauss(d, corr=True):
    """benchmarks.py:12-26 family: mu=0, C = 0.1 I + 0.9 11^T (corr) or I."""
    C = 0.1 * np.eye(d) + 0.9 * np.ones((d, d)) if corr else np.eye(d)
    return MultiGaussianDist(np.zeros(d), C)


class LogisticRegression(Model):
    """
    Bayesian logistic regression, prior N(0, prior_var I).  NOT IN THE REFERENCE
    (SURVEY.md section 8a row A11): follows the Model protocol of model.py:27-55.
    Parity unpinned by the reference; pinned analytically in tests.
    """

    def __init__(self, X, y, prior_var=100.0):
        self.X = np.asarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        if self.X.shape[0] != self.y.shape[0]:
            raise ParameterError("X and y have incompatible shapes")
        self.prior_var = float(prior_var)
        self.Ndim = self.X.shape[1]

    def log_prior(self, theta):
        d = self.Ndim
        return (-0.5 * np.dot(theta, theta) / self.prior_var
                - 0.5 * d * np.log(2 * np.pi * self.prior_var))

    def log_likelihood(self, theta):
        z = self.X @ theta
        return np.sum(self.y * z - np.logaddexp(0.0, z))

    def grad_log_posterior(self, theta):
        z = self.X @ theta
        return self.X.T @ (self.y - expit(z)) - theta / self.prior_var

    def metric(self, theta):
        """Fisher information + prior precision:  X^T diag(p(1-p)) X + I/prior_var."""
        p = expit(self.X @ theta)
        w = p * (1.0 - p)
        return (self.X.T * w) @ self.X + np.eye(self.Ndim) / self.prior_var


# --------------------------------------------------------------------------
# changepoint regression                    riemann/models/changepoint.py:22-240
# --------------------------------------------------------------------------
class ChangepointParams(object):
    """changepoint.py:22-45 -- (cpx[k], cpv[k+1], sig)."""

    def __init__(self, cpx, cpv, sig):
        if len(cpv) != len(cpx) + 1:                        # changepoint.py:35-39
            raise ValueError("number of constant pieces must be 1 more than "
                             "number of changepoints")
        self.cpx = np.array(cpx, dtype=np.float64)
        self.cpv = np.array(cpv, dtype=np.float64)
        self.sig = np.float64(np.squeeze(sig))      # numpy scalar: 1/0 -> inf like the reference

    def copy(self):
        return ChangepointParams(self.cpx, self.cpv, self.sig)


def birth_map_logjac(h, u):
    """log|det d(h1,h2)/d(h,u)| of changepoint.py:48-59 = log(h/(u(1-u)))."""
    return np.log(np.abs(h / (u * (1.0 - u))))


class ChangepointRegression1D(Model):
    """Piecewise-constant regression, changepoint.py:81-181 and :193-240."""

    def __init__(self, x, y, xmin, xmax, lamb, kmax, alpha, beta):
        if len(x) != len(y):                                # changepoint.py:91-95
            raise ValueError("length of predictor array must be same as response")
        self.x = np.array(x, dtype=np.float64)
        self.y = np.array(y, dtype=np.float64)
        self.xmin, self.xmax = float(xmin), float(xmax)
        self.kmax = kmax        # stored, never enforced (changepoint.py:100)
        self.lamb, self.alpha, self.beta = float(lamb), float(alpha), float(beta)

    def predict(self, theta, x):
        return theta.cpv[np.searchsorted(theta.cpx, x)]     # changepoint.py:181

    def log_likelihood(self, theta):
        # changepoint.py:106-126
        M = len(self.x)
        s2 = theta.sig ** 2
        with np.errstate(all="ignore"):
            r = self.predict(theta, self.x) - self.y
            logL = -0.5 * (np.sum(r * r / s2) + M * np.log(s2) + M * np.log(2 * np.pi))
        return -np.inf if np.isnan(logL) else logL

    def log_prior(self, theta):
        # changepoint.py:128-160; k is the number of STEPS in all three k-terms
        k = len(theta.cpv)
        a, b = self.alpha, self.beta
        with np.errstate(all="ignore"):
            lp_k = k * np.log(self.lamb) - gammaln(k) - self.lamb
            lp_v = np.sum(a * np.log(b) + (a - 1) * np.log(theta.cpv)
                          - b * theta.cpv - gammaln(a))
            s = np.concatenate([[self.xmin], theta.cpx, [self.xmax]])
            lp_s = (gammaln(2 * k + 1) + np.sum(np.log(s[1:] - s[:-1]))
                    - k * np.log(self.xmax - self.xmin))
            lp_sig = np.log(1.0 / theta.sig ** 2)
            if theta.sig < 0:
                lp_sig = np.nan
            lp = lp_k + lp_v + lp_s + lp_sig
        return -np.inf if np.isnan(lp) else lp

    def generate_synthetic_data(self, theta, Ndata):
        # changepoint.py:162-173 (draws: uniform(size=N) then normal(size=N))
        L = self.xmax - self.xmin
        x = np.sort(self.xmin + L * np.random.uniform(size=(Ndata,)))
        eps = np.random.normal(size=x.shape)
        return x, self.predict(theta, x) + theta.sig * eps

    def add_changepoint(self, theta, s, u):
        # changepoint.py:193-218
        if not (self.xmin < s < self.xmax):
            raise ValueError("require xmin < s < xmax for new changepoint")
        if not (0 < u < 1):
            raise ValueError("require 0 < u < 1 for new changepoint")
        n = int(np.searchsorted(theta.cpx, s))
        h = theta.cpv[n]
        f = np.sqrt((1 - u) / u)
        cpx = np.concatenate([theta.cpx[:n], [s], theta.cpx[n:]])
        cpv = np.concatenate([theta.cpv[:n], [h / f, h * f], theta.cpv[n + 1:]])
        return ChangepointParams(cpx, cpv, theta.sig), birth_map_logjac(h, u)

    def subtract_changepoint(self, theta, n):
        # changepoint.py:220-240
        if n not in range(len(theta.cpx)):
            raise ValueError("n has to be a valid changepoint index")
        h1, h2 = theta.cpv[n], theta.cpv[n + 1]
        h = np.sqrt(h1 * h2)
        u = 1.0 / (1 + h2 / h1)
        cpx = np.concatenate([theta.cpx[:n], theta.cpx[n + 1:]])
        cpv = np.concatenate([theta.cpv[:n], [h], theta.cpv[n + 2:]])
        return ChangepointParams(cpx, cpv, theta.sig), -birth_map_logjac(h, u)


# --------------------------------------------------------------------------
# Proposal protocol                       riemann/proposals/proposal.py:1-26
# --------------------------------------------------------------------------
class Proposal(object):
    draws = LiveDraws()

    def propose(self, theta):
        raise NotImplementedError("abstract")

    def adapt(self, theta):
        pass


class MetropolisRandomWalk(Proposal):
    """theta + scale * L xi, logqratio 0.       riemann/proposals/randomwalk.py:12-26"""

    def __init__(self, C):
        self.scale = 1.0
        self.L = np.linalg.cholesky(np.atleast_2d(C))

    def propose(self, theta):
        theta = np.atleast_1d(theta)
        if self.L.shape[1] != theta.shape[0]:               # randomwalk.py:23-24
            raise ParameterError("theta and L have incompatible shapes")
        xi = self.draws.normal("xi", theta.shape[0])
        return theta + self.scale * np.dot(self.L, xi), 0.0


class AdaptScaleProposal(Proposal):
    """Acceptance-rate-targeting scale.      riemann/proposals/adaptive.py:11-35"""

    def __init__(self, target_accept_rate):
        self.Nsamples = 0
        self.Naccepts = 0
        self.last_theta = None
        self.accept_rate = 0.0
        self.target_accept_rate = target_accept_rate
        self.scale = 1.0

    def adapt(self, theta):
        self.Nsamples += 1
        # adaptive.py:28 -- "accept" == state changed; first call compares with None
        moved = True if self.last_theta is None else bool(np.any(theta != self.last_theta))
        self.Naccepts += moved
        self.accept_rate = self.Naccepts / float(self.Nsamples)
        r = np.exp(1.0 / float(self.Nsamples))
        if self.accept_rate > self.target_accept_rate:
            self.scale *= r
        else:
            self.scale /= r
        self.last_theta = theta


class AdaptScaleRandomWalk(AdaptScaleProposal, MetropolisRandomWalk):
    """randomwalk.py:29-37, target 0.25."""

    def __init__(self, C):
        AdaptScaleProposal.__init__(self, 0.25)
        MetropolisRandomWalk.__init__(self, C)


class AdaptCovProposal(Proposal):
    """adaptive.py:38-103 (Haario et al. 2001): C follows the sample covariance of the chain history, recomputed when
    `np.sqrt(n)**2 == n and n > 2` (the perfect squares and, through rounding, about half of the other n).
    The strict (non-smooth) mode with n < t_adapt is not restated: there the reference divides its current C in
    place (adaptive.py:89-91 stores to a dead attribute, :101 then rescales the old C, which on the first occasion
    is the caller's C0 array itself); t_adapt <= 4 never gets there."""

    def __init__(self, C0, t_adapt=1, marginalize=False, smooth_adapt=False):
        if not smooth_adapt and t_adapt > 4:
            raise ParameterError("strict Haario mode with t_adapt > 4 is an in-place aliasing bug in the reference")
        self.C0 = np.array(np.atleast_2d(C0), dtype=np.float64)
        self.C = self.C0
        self.L = np.linalg.cholesky(self.C0)
        self.t_adapt, self.marginalize, self.smooth_adapt = t_adapt, marginalize, smooth_adapt
        self._S, self._SX, self._SX2 = 0.0, 0.0, 0.0

    def adapt(self, theta):
        X = np.atleast_1d(theta)
        self._S += 1
        self._SX = self._SX + X
        self._SX2 = self._SX2 + X[:, None] * X[None, :]
        n = float(self._S)
        # adaptive.py:82-83.  NB: on numpy scalars `x ** 2` is libm pow(x, 2.0); the test is true for all perfect squares
        # and, through rounding, for about half of the other n (and not for the same n as `x * x == n`: 238, 952, ...)
        if np.sqrt(n) ** 2 == n and n > 2:
            Cs = (self._SX2 - self._SX[:, None] * self._SX[None, :] / n) / (n - 1)
            if self.smooth_adapt:
                C = (n * Cs + self.t_adapt * self.C0) / (n + self.t_adapt)        # :86-87
            else:
                Creg = np.mean(np.diag(self.C0)) * np.eye(Cs.shape[0])
                C = Cs + 1e-12 * Creg                                             # :92-93
            if self.marginalize:
                C = np.diag(np.diag(C))                                           # :96-97
            d = C.shape[0]
            self.C = C / d ** 0.4                                                 # :101
            self.L = np.linalg.cholesky(self.C) / d ** 0.2                        # :102


class AdaptCovRandomWalk(AdaptCovProposal, MetropolisRandomWalk):
    """randomwalk.py:40-56."""

    def __init__(self, C0, t_adapt=1, marginalize=False, smooth_adapt=False):
        MetropolisRandomWalk.__init__(self, C0)
        AdaptCovProposal.__init__(self, C0, t_adapt=t_adapt, marginalize=marginalize, smooth_adapt=smooth_adapt)


class pCN(Proposal):
    """Preconditioned Crank-Nicolson.        riemann/proposals/randomwalk.py:78-100"""

    def __init__(self, C, rho):
        self.scale = 1.0
        self.rho = rho
        self.rho_c = np.sqrt(1 - rho ** 2)
        self.L = np.linalg.cholesky(np.atleast_2d(C))

    def propose(self, theta):
        theta = np.atleast_1d(theta)
        xi = self.draws.normal("xi", theta.shape[0])
        theta_p = self.rho * theta + self.rho_c * np.dot(self.L, xi)
        Ls = self.L * self.rho_c
        u_fwd = solve_triangular(Ls, theta_p - self.rho * theta, lower=True, check_finite=False)
        u_rev = solve_triangular(Ls, theta - self.rho * theta_p, lower=True, check_finite=False)
        return theta_p, -0.5 * (np.dot(u_fwd, u_fwd) - np.dot(u_rev, u_rev))


class AdaptScalepCN(AdaptScaleProposal, pCN):
    """randomwalk.py:103-119, target 0.25.  As written there: rho is re-derived from the PREVIOUS rho at every
    proposal (`rho = tanh(rho / scale)`, compounding -- SURVEY appendix B) and rho_c keeps its initial value."""

    def __init__(self, C, rho):
        AdaptScaleProposal.__init__(self, 0.25)
        pCN.__init__(self, C, rho)
        self.rho0 = self.rho

    def propose(self, theta):
        self.rho = np.tanh(self.rho / self.scale)
        return pCN.propose(self, theta)


class AdaptScaleCovRandomWalk(AdaptCovRandomWalk):
    """randomwalk.py:62-75: scale adaptation (target 0.25) followed by covariance adaptation, every step."""

    def __init__(self, C0, t_adapt=1, marginalize=False, smooth_adapt=False):
        AdaptCovRandomWalk.__init__(self, C0, t_adapt=t_adapt, marginalize=marginalize, smooth_adapt=smooth_adapt)
        self._sc = AdaptScaleProposal(0.25)
        self.scale = 1.0

    def adapt(self, theta):
        self._sc.scale = self.scale
        self._sc.adapt(theta)                                   # AdaptScaleRandomWalk.adapt  (:73)
        self.scale, self.accept_rate = self._sc.scale, self._sc.accept_rate
        AdaptCovRandomWalk.adapt(self, theta)                   # :74


def leapfrog(p0, q0, Nsteps, eps, grad, M=None):
    """Stormer-Verlet.                  riemann/proposals/hamiltonian.py:13-52"""
    vel = (lambda p: np.linalg.solve(M, p)) if M is not None else (lambda p: p)
    p = p0 + 0.5 * eps * grad(q0)                           # :27
    q = q0 + eps * vel(p)                                   # :29-30
    for _ in range(Nsteps - 1):                             # :34-38
        p = p + eps * grad(q)
        q = q + eps * vel(p)
    p = p + 0.5 * eps * grad(q)                             # :40
    return p, q


class VanillaHMC(Proposal):
    """riemann/proposals/hamiltonian.py:55-91.  Nsteps=1 is (preconditioned) MALA."""

    def __init__(self, eps, Nsteps, gradlogpost, M=None):
        self.Nsteps = Nsteps
        self.eps = eps
        self._grad = gradlogpost
        self.M = M
        self.chM = None if M is None else np.linalg.cholesky(M)

    def propose(self, theta):
        theta = np.atleast_1d(theta)
        p0 = self.draws.normal("xi", theta.shape[0])                  # :79
        if self.chM is not None:
            p0 = np.dot(self.chM, p0)                                  # :81
        p1, theta_new = leapfrog(p0, theta, self.Nsteps, self.eps, self._grad, self.M)
        if self.chM is not None:                                       # :85-87
            p0 = solve_triangular(self.chM, p0, lower=True, check_finite=False)
            p1 = solve_triangular(self.chM, p1, lower=True, check_finite=False)
        return theta_new, 0.5 * (np.sum(p1 ** 2) - np.sum(p0 ** 2))   # :89


class AdaptScaleHMC(AdaptScaleProposal, VanillaHMC):
    """hamiltonian.py:94-103, target 0.75, eps = scale * eps0."""

    def __init__(self, eps, Nsteps, gradlogpost, M=None):
        AdaptScaleProposal.__init__(self, 0.75)
        VanillaHMC.__init__(self, eps, Nsteps, gradlogpost, M=M)
        self.eps0 = self.eps

    def propose(self, theta):
        self.eps = self.scale * self.eps0
        return VanillaHMC.propose(self, theta)


class AdaptScaleCovHMC(AdaptScaleHMC):
    """hamiltonian.py:121-135.  Its `adapt` resolves to AdaptScaleProposal.adapt (MRO: AdaptScaleCovHMC, AdaptScaleHMC,
    AdaptScaleProposal, AdaptCovHMC, AdaptCovProposal, ...), which does not chain to AdaptCovProposal.adapt: the
    covariance is never adapted and the class behaves as AdaptScaleHMC with the fixed mass matrix M0 (probed: _S stays 0)."""

    def __init__(self, eps, Nsteps, gradlogpost, M0, t_adapt=1, marginalize=False, smooth_adapt=False):
        AdaptScaleHMC.__init__(self, eps, Nsteps, gradlogpost, M=np.array(M0, dtype=np.float64))
        self.C0 = self.C = self.M
        self.L = self.chM


class AdaptCovHMC(AdaptCovProposal, VanillaHMC):
    """hamiltonian.py:106-119: the leapfrog's mass matrix is the adapted covariance, M = C and chM = L -- where
    L = chol(C) / d**0.2 (adaptive.py:102), so after the first adaptation chM chM^T = M / d**0.4, as in the reference."""

    def __init__(self, eps, Nsteps, gradlogpost, M0, t_adapt=1, marginalize=False, smooth_adapt=False):
        VanillaHMC.__init__(self, eps, Nsteps, gradlogpost, M=np.array(M0, dtype=np.float64))
        AdaptCovProposal.__init__(self, M0, t_adapt=t_adapt, marginalize=marginalize, smooth_adapt=smooth_adapt)

    def propose(self, theta):
        self.M, self.chM = self.C, self.L
        return VanillaHMC.propose(self, theta)


def MALA(eps, gradlogpost, M=None):
    """MALA with h = eps^2 is exactly VanillaHMC(eps, 1, grad)."""
    return VanillaHMC(eps, 1, gradlogpost, M=M)


class SimplifiedMMALA(Proposal):
    """
    Simplified manifold MALA (Girolami & Calderhead 2011 sec. 5.3, metric-derivative
    terms dropped).  NOT IN THE REFERENCE -- follows the Proposal
    protocol (proposal.py:10-17) and the sign convention of sampler.py:83.

        G = model.metric(theta) = L L^T
        mean(theta) = theta + eps^2/2 * G^{-1} grad(theta)
        theta' = mean(theta) + eps * L^{-T} xi                   cov = eps^2 G^{-1}
        log q(b|a) = 1/2 logdet G(a) - d/2 log(2 pi eps^2) - |L(a)^T (b - mean(a))|^2 / (2 eps^2)
        logqratio = log q(theta'|theta) - log q(theta|theta')
    """

    def __init__(self, eps, model):
        self.eps = eps
        self.model = model

    def _geometry(self, theta):
        G = self.model.metric(theta)
        L = np.linalg.cholesky(G)
        g = self.model.grad_log_posterior(theta)
        nat = solve_triangular(L.T, solve_triangular(L, g, lower=True, check_finite=False), lower=False, check_finite=False)
        mean = theta + 0.5 * self.eps ** 2 * nat
        return L, mean, 2.0 * np.sum(np.log(np.diag(L)))

    def _logq(self, L, mean, logdet, b):
        r = np.dot(L.T, b - mean)
        d = len(b)
        return (0.5 * logdet - 0.5 * d * np.log(2 * np.pi * self.eps ** 2)
                - 0.5 * np.dot(r, r) / self.eps ** 2)

    def propose(self, theta):
        theta = np.atleast_1d(theta)
        xi = self.draws.normal("xi", theta.shape[0])
        L, mean, logdet = self._geometry(theta)
        theta_p = mean + self.eps * solve_triangular(L.T, xi, lower=False, check_finite=False)
        Lp, mean_p, logdet_p = self._geometry(theta_p)
        lqr = self._logq(L, mean, logdet, theta_p) - self._logq(Lp, mean_p, logdet_p, theta)
        return theta_p, lqr


class ChangepointRegression1DProp(Proposal):
    """
    The 4-way mixture proposal config 2 uses.   examples/test_changepoint.py:18-73
    Block selection consumes a FRESH uniform per elif (:48,51,54).
    """

    def __init__(self, model, hscale):
        self.model = model
        self.Ndata = len(model.x)
        self.hscale = hscale
        self.k = None

    def step_sizes(self, k):
        """sqrt of the isotropic variances at test_changepoint.py:36-38; like the reference (:45-46, `self.k`) the
        three Cholesky factors are rebuilt only when the number of changepoints has changed."""
        if k == self.k:
            return self._steps
        self.k, self._steps = k, self._step_sizes(k)
        return self._steps

    def _step_sizes(self, k):
        xmin, xmax = self.model.xmin, self.model.xmax
        sx = np.linalg.cholesky(0.01 * (xmax - xmin) / (k + 1) * np.eye(max(k, 1)))[0, 0]
        sv = np.linalg.cholesky(0.01 * self.hscale ** 2 / self.Ndata * np.eye(k + 1))[0, 0]
        ss = np.linalg.cholesky(np.atleast_2d(0.01 * self.hscale))[0, 0]
        return sx, sv, ss

    def propose(self, theta):
        dr = self.draws
        k = len(theta.cpx)
        sx, sv, ss = self.step_sizes(k)
        new = theta.copy()
        if dr.uniform("sel1") < 0.20:
            new.cpx = theta.cpx + 1.0 * (sx * dr.normal("xi", k))
            return new, 0.0
        if dr.uniform("sel2") < 0.40:
            new.cpv = theta.cpv + 1.0 * (sv * dr.normal("xi", k + 1))
            return new, 0.0
        if dr.uniform("sel3") < 0.60:
            new.sig = np.float64(theta.sig + 1.0 * (ss * dr.normal("xi", 1)[0]))
            return new, 0.0
        if k == 0 or dr.uniform("bd") > 0.5:                            # :59
            s = dr.uniform("s", self.model.xmin, self.model.xmax)
            u = 0.5 + dr.uniform("du", -0.1, 0.1) / np.sqrt(self.Ndata)
            return self.model.add_changepoint(theta, s, u)
        return self.model.subtract_changepoint(theta, dr.randint("n", k))


# --------------------------------------------------------------------------
# Sampler                                     riemann/samplers/sampler.py:28-90
# --------------------------------------------------------------------------
class Sampler(object):
    def __init__(self, model, proposal, theta0, draws=None):
        self.model = model
        self.proposal = proposal
        self.draws = draws if draws is not None else LiveDraws()
        self.proposal.draws = self.draws
        self._chain_thetas = [theta0]
        self._chain_logpost = [model.log_posterior(theta0)]
        self.last_proposal = None       # (theta', logpost', logqratio, accepted) -- test hook

    def run(self, Nsamples, Nburn=0, Nthin=1):
        self._chain_thetas = self._chain_thetas[-1:]        # sampler.py:49-50 (resume)
        self._chain_logpost = self._chain_logpost[-1:]
        for _ in range(Nsamples):
            self.sample()
        self._chain_thetas = self._chain_thetas[Nburn::Nthin]   # :53-54
        self._chain_logpost = self._chain_logpost[Nburn::Nthin]

    def current_state(self):
        return self._chain_thetas[-1], self._chain_logpost[-1]

    def _add_state(self, theta, logpost):
        self._chain_thetas.append(theta)
        self._chain_logpost.append(logpost)

    def sample(self):
        theta_old, lp_old = self.current_state()
        theta_prop, lqr = self.proposal.propose(theta_old)
        lp_prop = self.model.log_posterior(theta_prop)
        with np.errstate(invalid="ignore"):
            delta = lp_prop - lp_old - lqr
        mh = delta if delta < 0 else 0           # Python min(0, x): nan -> 0  (sampler.py:83)
        accepted = bool(np.log(self.draws.uniform("acc")) < mh)          # strict <, :84
        theta, lp = (theta_prop, lp_prop) if accepted else (theta_old, lp_old)
        self.last_proposal = (theta_prop, lp_prop, lqr, accepted)
        self.proposal.adapt(theta)                                        # :88
        self._add_state(theta, lp)
        self.draws.next_step()
        return theta, lp


# --------------------------------------------------------------------------
# Parallel tempering ("next" row N3)               riemann/samplers/ptsampler.py:11-127
# --------------------------------------------------------------------------
class PTTapeDraws(object):
    """Replay for PTSampler: usel[T, Nt] selection uniforms, xi[T, Nt, d] normals and u[T, Nt] accept /
    swap uniforms, addressed by the chain the draw is made for."""

    def __init__(self, usel, xi, u):
        self.usel, self.xi, self.u = (np.asarray(a, dtype=np.float64) for a in (usel, xi, u))
        self.t, self.i = 0, 0

    def normal(self, role, n):
        return self.xi[self.t, self.i, :n].copy()

    def uniform(self, role, low=0.0, high=1.0):
        return self.usel[self.t, self.i] if role == "ptsel" else self.u[self.t, self.i]

    def randint(self, role, n):
        raise NotImplementedError

    def next_step(self):
        pass                                    # PTSampler advances t itself (one PT step = Nt chains)


class TemperedModel(Model):
    """ptsampler.py:11-38: likelihood scaled by beta, prior untouched."""

    def __init__(self, base_model, beta):
        if not (beta >= 0 and beta <= 1):
            raise ParameterError("beta = {} must be a number between 0 and 1".format(beta))
        self.base_model, self.beta = base_model, beta

    def log_likelihood(self, theta):
        return self.base_model.log_likelihood(theta) * self.beta

    def log_prior(self, theta):
        return self.base_model.log_prior(theta)


class PTSampler(object):
    """ptsampler.py:41-127 with the default ladder betas = 0.5**arange(5) (the `betas=` branch of the
    reference raises: `isinstance(betas, np.array)`, :62); an explicit ladder is accepted here and used
    as given, which is what that branch intends (:66-68 computes a sorted copy and then ignores it)."""

    def __init__(self, model, proposal, theta0, betas=None, Pswap=0.1, draws=None):
        self.betas = 0.5 ** np.arange(5) if betas is None else np.asarray(betas, dtype=np.float64)
        if not (Pswap > 0 and Pswap < 1):
            raise ParameterError("Pswap must be a number between 0 and 1")
        self.Pswap = Pswap
        self.draws = draws if draws is not None else LiveDraws()
        self.samplers = [Sampler(TemperedModel(model, b), proposal, theta0, draws=self.draws) for b in self.betas]

    def run(self, Nsamples, Nburn=0, Nthin=1):
        for _ in range(Nsamples):                               # :83-90: Nburn / Nthin are ignored
            self.sample()
        self._chain_thetas = self.samplers[0]._chain_thetas
        self._chain_logpost = self.samplers[0]._chain_logpost

    def sample(self):
        swapped = []
        n = len(self.samplers)
        tape = isinstance(self.draws, PTTapeDraws)
        for i in range(n):
            if tape:
                self.draws.i = i
            u = self.draws.uniform("ptsel")                     # :104 drawn for EVERY chain, swapped or not
            if i in swapped:
                continue
            elif u > self.Pswap or i == n - 1:
                self.samplers[i].sample()                       # :110
            else:
                j = i + 1
                th_i, lp_ii = self.samplers[i].current_state()
                th_j, lp_jj = self.samplers[j].current_state()
                lp_ij = self.samplers[i].model.log_posterior(th_j)
                lp_ji = self.samplers[j].model.log_posterior(th_i)
                with np.errstate(all="ignore"):
                    x = np.exp((lp_ji + lp_ij) - (lp_ii + lp_jj))
                mh = x if x < 1 else 1                          # Python min(1, x): nan -> 1  (:119-120)
                if self.draws.uniform("acc") < mh:              # :121
                    self.samplers[i]._add_state(th_j, lp_ij)
                    self.samplers[j]._add_state(th_i, lp_ji)
                else:
                    self.samplers[i]._add_state(th_i, lp_ii)
                    self.samplers[j]._add_state(th_j, lp_jj)
                swapped.extend([i, j])
        if tape:
            self.draws.t += 1


# --------------------------------------------------------------------------
# synthetic inputs of SURVEY.md section 8d (numpy Philox generator => reproducible)
# --------------------------------------------------------------------------
SEED_BASE = 20261018


def make_changepoint_problem(seed=SEED_BASE + 2, Ncpx=5, Ndata=100, xmin=1.0, xmax=3.0,
                             hmin=1.0, hmax=3.0, sig=0.1):
    """Recipe of examples/test_changepoint.py:137-150,167 with a seeded generator: the arrays of
    riemann_b200/synthetic.py (plain numpy, the one source of the bench inputs) wrapped in the port's classes."""
    from riemann_b200 import synthetic
    c = synthetic.changepoint_problem(seed, Ncpx, Ndata, xmin, xmax, hmin, hmax, sig)
    model = ChangepointRegression1D(c["x"], c["y"], c["xmin"], c["xmax"], c["lamb"], c["kmax"], c["alpha"], c["beta"])
    theta0 = ChangepointParams(list(c["theta0"][0]), list(c["theta0"][1]), c["theta0"][2])
    return model, ChangepointRegression1DProp(model, c["hscale"]), theta0, ChangepointParams(*c["theta_true"])


def make_logistic_problem(N, d, seed=SEED_BASE + 4, prior_var=100.0, dtype=np.float64):
    from riemann_b200 import synthetic
    return synthetic.logistic_problem(N, d, seed, prior_var, dtype)
