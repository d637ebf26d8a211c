"""HMC / MALA proposals -- riemann/proposals/hamiltonian.py:13-103."""
import ctypes as C

import numpy as np

from .. import _lib
from ..models.model import DeviceModel
from ..sampling_errors import ParameterError
from .adaptive import AdaptScaleProposal
from .proposal import DeviceProposal


class VanillaHMC(DeviceProposal):
    """
    hamiltonian.py:55-91.  ``Nsteps = 1`` is MALA with h = eps^2;
    with ``M`` it is preconditioned MALA with covariance M^-1.

    ``gradlogpost`` keeps the reference's position in the signature but must be the
    gradient method of a device model (e.g. ``model.grad_log_likelihood`` as in
    examples/test_randomwalk.py:32): the engine evaluates that model's gradient
    in-kernel and caches it across iterations (the reference recomputes it twice per
    step, hamiltonian.py:27,40).  An arbitrary Python callable has no device kernel and
    is rejected -- there is no CPU fallback.
    """

    _adaptive = False

    def __init__(self, eps, Nsteps, gradlogpost, M=None):
        self.Nsteps = Nsteps
        self.eps = eps
        self._mgradU = gradlogpost
        owner = getattr(gradlogpost, "__self__", None)
        if not isinstance(owner, DeviceModel):
            raise ParameterError(
                "gradlogpost must be a gradient method of a riemann_b200 device model "
                "(no device kernel exists for an arbitrary Python callable)")
        self._model = owner
        if M is None:
            self.M, self.chM = None, None
        else:
            self.M = np.atleast_2d(np.asarray(M, dtype=np.float64))
            try:
                self.chM = np.linalg.cholesky(self.M)                   # hamiltonian.py:74
            except np.linalg.LinAlgError as e:
                raise ParameterError("mass matrix is not positive definite: {}".format(e))

    def _create_handle(self, d):
        h = C.c_void_p()
        if self.M is not None:
            if self.M.shape != (d, d):
                raise ParameterError("theta and M have incompatible shapes")
            chM = np.ascontiguousarray(self.chM)
            Minv = np.ascontiguousarray(np.linalg.inv(self.M))
            chMinv = np.ascontiguousarray(np.linalg.solve(self.chM, np.eye(d)))
            args = (_lib.ptr(chM), _lib.ptr(Minv), _lib.ptr(chMinv))
        else:
            args = (None, None, None)
        eps0 = float(getattr(self, "eps0", self.eps))
        _lib.check(_lib.load().rmn_proposal_hmc_create(
            C.byref(h), d, eps0, int(self.Nsteps), args[0], args[1], args[2],
            1 if self._adaptive else 0, float(getattr(self, "target_accept_rate", 0.75))))
        return h


class AdaptScaleHMC(AdaptScaleProposal, VanillaHMC):
    """hamiltonian.py:94-103: eps = scale * eps0, target acceptance 0.75."""

    def __init__(self, eps, Nsteps, gradlogpost, M=None):
        AdaptScaleProposal.__init__(self, 0.75)
        VanillaHMC.__init__(self, eps, Nsteps, gradlogpost, M=M)
        self.eps0 = self.eps


class AdaptCovHMC(VanillaHMC):
    """hamiltonian.py:106-119: HMC whose mass matrix is the chain's adapted covariance (adaptive.py:38-103), M = C and
    chM = L -- L = chol(C) / d**0.2 there, so chM chM^T = M / d**0.4 after the first adaptation; reproduced as written.
    Per chain on the device; `.C` / `.L` / `.M` / `.chM` refreshed after every batch.  Small-d path (d <= 8)."""

    _adapt_cov = True

    def __init__(self, eps, Nsteps, gradlogpost, M0, t_adapt=1, marginalize=False, smooth_adapt=False):
        VanillaHMC.__init__(self, eps, Nsteps, gradlogpost, M=None)
        self.C0 = np.array(np.atleast_2d(M0), dtype=np.float64)
        self.C = self.M = self.C0
        self.L = self.chM = np.linalg.cholesky(self.C0)
        self.t_adapt, self.marginalize, self.smooth_adapt = t_adapt, marginalize, smooth_adapt

    def _create_handle(self, d):
        if self.C0.shape != (d, d):
            raise ParameterError("theta and M0 have incompatible shapes")
        M_save, chM_save = self.M, self.chM
        self.M = self.chM = None                              # created WITHOUT a mass matrix; the adapted one takes its place
        try:
            h = VanillaHMC._create_handle(self, d)
        finally:
            self.M, self.chM = M_save, chM_save
        C0 = np.ascontiguousarray(self.C0)
        L0 = np.ascontiguousarray(np.linalg.cholesky(self.C0))
        _lib.check(_lib.load().rmn_proposal_hmc_set_cov_adapt(h, _lib.ptr(C0), _lib.ptr(L0), float(self.t_adapt),
                                                              int(bool(self.marginalize)), int(bool(self.smooth_adapt))))
        return h


class AdaptScaleCovHMC(AdaptScaleHMC):
    """hamiltonian.py:121-135 -- what examples/test_randomwalk.py:36-38 (BASELINE config 0) runs.  In the reference its
    `adapt` resolves to AdaptScaleProposal.adapt (MRO: AdaptScaleCovHMC, AdaptScaleHMC, AdaptScaleProposal, AdaptCovHMC,
    AdaptCovProposal, ...), which never chains to AdaptCovProposal.adapt: the mass matrix stays M0 for the whole run
    (probed on the reference: `_S` stays 0).  So this is AdaptScaleHMC with the fixed mass matrix M0; `t_adapt`,
    `marginalize` and `smooth_adapt` are accepted and, like there, have no effect."""

    def __init__(self, eps, Nsteps, gradlogpost, M0, t_adapt=1, marginalize=False, smooth_adapt=False):
        AdaptScaleHMC.__init__(self, eps, Nsteps, gradlogpost, M=np.array(M0, dtype=np.float64))
        self.C0 = self.C = np.array(M0, dtype=np.float64)
        self.L = np.linalg.cholesky(self.C)
        self.t_adapt, self.marginalize, self.smooth_adapt = t_adapt, marginalize, smooth_adapt


def MALA(eps, gradlogpost, M=None):
    """Metropolis-adjusted Langevin == VanillaHMC(eps, 1, grad[, M])."""
    return VanillaHMC(eps, 1, gradlogpost, M=M)


class SimplifiedMMALA(DeviceProposal):
    """
    Simplified manifold MALA with the model's Fisher metric G(theta) (not in the
    reference; Proposal protocol proposal.py:10-17):
        theta' ~ N(theta + eps^2/2 G^-1 grad, eps^2 G^-1),
        logqratio = log q(theta'|theta) - log q(theta|theta') including the logdet terms.
    """

    _adaptive = False

    def __init__(self, eps, model):
        if not isinstance(model, DeviceModel) or not hasattr(model, "metric_batch"):
            raise ParameterError("SimplifiedMMALA needs a device model with a Fisher metric")
        self.eps = eps
        self._model = model

    def _create_handle(self, d):
        h = C.c_void_p()
        _lib.check(_lib.load().rmn_proposal_mmala_create(C.byref(h), d, float(self.eps)))
        return h
