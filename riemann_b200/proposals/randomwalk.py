"""Random-walk proposals -- riemann/proposals/randomwalk.py:12-37, 78-100."""
import ctypes as C

import numpy as np

from .. import _lib
from ..sampling_errors import ParameterError
from .adaptive import AdaptScaleProposal
from .proposal import DeviceProposal


def _chol(C_):
    try:
        return np.linalg.cholesky(np.atleast_2d(np.asarray(C_, dtype=np.float64)))
    except np.linalg.LinAlgError as e:
        raise ParameterError("proposal covariance is not positive definite: {}".format(e))


class MetropolisRandomWalk(DeviceProposal):
    """theta' = theta + scale * L xi,  xi ~ N(0, I),  logqratio = 0  (randomwalk.py:17-26)."""

    _adaptive = False

    def __init__(self, C_):
        self.scale = 1.0
        self.L = _chol(C_)

    def _create_handle(self, d):
        if self.L.shape[1] != d:                                        # randomwalk.py:23-24
            raise ParameterError("theta and L have incompatible shapes")
        h = C.c_void_p()
        L = np.ascontiguousarray(self.L)
        _lib.check(_lib.load().rmn_proposal_rw_create(
            C.byref(h), d, _lib.ptr(L), 1 if self._adaptive else 0,
            float(getattr(self, "target_accept_rate", 0.25))))
        return h


class AdaptScaleRandomWalk(AdaptScaleProposal, MetropolisRandomWalk):
    """randomwalk.py:29-37: adapts the overall scale towards 25 % acceptance."""

    def __init__(self, C_):
        AdaptScaleProposal.__init__(self, 0.25)
        MetropolisRandomWalk.__init__(self, C_)


class AdaptCovRandomWalk(MetropolisRandomWalk):
    """randomwalk.py:40-56 + adaptive.py:38-103 (Haario et al. 2001): the proposal covariance follows the sample
    covariance of the chain history; per chain, adapted and refactored inside the device kernel.  `.C` / `.L` are
    refreshed from the device after every batch (arrays [K, d, d] for K > 1 chains).  Small-d path (d <= 8)."""

    _adaptive = False
    _adapt_cov = True

    def __init__(self, C0, t_adapt=1, marginalize=False, smooth_adapt=False):
        MetropolisRandomWalk.__init__(self, C0)
        self.C0 = np.array(np.atleast_2d(C0), dtype=np.float64)
        self.C = self.C0
        self.t_adapt, self.marginalize, self.smooth_adapt = t_adapt, marginalize, smooth_adapt

    def _create_handle(self, d):
        if self.L.shape[1] != d:
            raise ParameterError("theta and L have incompatible shapes")
        h = C.c_void_p()
        C0 = np.ascontiguousarray(self.C0)
        L0 = np.ascontiguousarray(np.linalg.cholesky(self.C0))
        _lib.check(_lib.load().rmn_proposal_adaptcov_create(C.byref(h), d, _lib.ptr(C0), _lib.ptr(L0),
                                                            float(self.t_adapt), int(bool(self.marginalize)),
                                                            int(bool(self.smooth_adapt))))
        return h


class PooledAdaptCovRandomWalk(MetropolisRandomWalk):
    """Covariance adaptation POOLED over the chains.  Haario's adaptive Metropolis as the reference
    runs it (adaptive.py:38-103) learns the proposal covariance from ONE chain's history; a K-chain engine has a better
    estimator at hand, the population: every `t_adapt` steps the current states of all chains (all ranks, when
    torch.distributed is initialised) join running sums, and from the next step on every chain proposes with
    `C = sd * Cov(pool) + jitter * I`, `sd = 2.38^2 / d` by default (Haario et al. 2001).  `stop_after` > 0 freezes the
    proposal after that many steps (adaptation during burn-in only, after which the chains are plain MH chains);
    0 adapts forever with ever smaller changes.  Dense Gaussian path (d > 8), precision f64.  After every batch `.C`
    (the pooled covariance estimate), `.pool_mean` and `.pool_count` are refreshed from the device.
    `adapt_scale=True` also runs the per-chain AdaptScale rule (adaptive.py:26-35, target 0.25) on top."""

    _adaptive = False
    _pooled_cov = True

    def __init__(self, C0, t_adapt=100, sd=None, jitter=1e-10, stop_after=0, adapt_scale=False):
        MetropolisRandomWalk.__init__(self, C0)
        if not (int(t_adapt) >= 1):
            raise ParameterError("t_adapt must be a positive number of steps")
        self.C0 = np.array(np.atleast_2d(C0), dtype=np.float64)
        self.C, self.pool_mean, self.pool_count = None, None, 0.0
        self.t_adapt, self.sd, self.jitter, self.stop_after = int(t_adapt), sd, float(jitter), int(stop_after)
        if adapt_scale:
            self._adaptive = True
            self.target_accept_rate = 0.25
            self.accept_rate = 0.0

    def _create_handle(self, d):
        h = MetropolisRandomWalk._create_handle(self, d)
        sd = 2.38 ** 2 / d if self.sd is None else float(self.sd)
        _lib.check(_lib.load().rmn_proposal_rw_set_pooled_cov_adapt(h, self.t_adapt, sd, self.jitter, self.stop_after))
        return h


# aliases of the reference (randomwalk.py:59)
AdaptiveMetropolisRandomWalk = HaarioRandomWalk = AdaptCovRandomWalk


class AdaptScaleCovRandomWalk(AdaptScaleProposal, AdaptCovRandomWalk):
    """randomwalk.py:62-75: adapts both the scale (target 0.25) and the covariance, in that order, every step."""

    def __init__(self, C0, t_adapt=1, marginalize=False, smooth_adapt=False):
        AdaptScaleProposal.__init__(self, 0.25)
        AdaptCovRandomWalk.__init__(self, C0, t_adapt=t_adapt, marginalize=marginalize, smooth_adapt=smooth_adapt)
        self._adaptive = True

    def _create_handle(self, d):
        h = AdaptCovRandomWalk._create_handle(self, d)
        _lib.check(_lib.load().rmn_proposal_set_scale_adapt(h, 1, float(self.target_accept_rate)))
        return h


class pCN(DeviceProposal):
    """Preconditioned Crank-Nicolson (randomwalk.py:78-100)."""

    _adaptive = False

    def __init__(self, C_, rho):
        self.scale = 1.0
        self.rho = rho
        self.rho_c = np.sqrt(1 - self.rho ** 2)
        self.L = _chol(C_)

    def _create_handle(self, d):
        if self.L.shape[1] != d:
            raise ParameterError("theta and L have incompatible shapes")
        h = C.c_void_p()
        L = np.ascontiguousarray(self.L)
        Linv = np.ascontiguousarray(np.linalg.solve(self.L, np.eye(d)))
        _lib.check(_lib.load().rmn_proposal_pcn_create(C.byref(h), d, _lib.ptr(L), _lib.ptr(Linv),
                                                       float(self.rho)))
        return h


class AdaptScalepCN(AdaptScaleProposal, pCN):
    """randomwalk.py:103-119, target 0.25 -- reproduced as written: every proposal re-derives rho from the previous
    rho (`rho = tanh(rho / scale)`, compounding) while rho_c keeps its initial value.  Per chain on
    the device; `.rho` is not refreshed on the host.  Small-d path (d <= 8)."""

    def __init__(self, C_, rho):
        AdaptScaleProposal.__init__(self, 0.25)
        pCN.__init__(self, C_, rho)
        self.rho0 = self.rho
        self._adaptive = True

    def _create_handle(self, d):
        h = pCN._create_handle(self, d)
        _lib.check(_lib.load().rmn_proposal_set_scale_adapt(h, 1, float(self.target_accept_rate)))
        return h
