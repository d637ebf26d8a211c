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
