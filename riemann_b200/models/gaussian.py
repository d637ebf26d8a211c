"""MultiGaussianDist -- riemann/models/gaussian.py:21-66 on the device."""
import ctypes as C

import numpy as np

from .. import _lib
from ..sampling_errors import ParameterError
from .model import DeviceModel


class MultiGaussianDist(DeviceModel):
    """
    Multivariate Gaussian with known mean and covariance (gaussian.py:21-58).
    Host side (setup only, fp64 numpy): L = chol(C), logdetC, P = C^-1, L^-1.
    Device side: log-density and gradient from one product  v = P (theta - mu):
    grad = -v, quadratic form = (theta - mu) . v.
    """

    def __init__(self, mu, C_):
        DeviceModel.__init__(self)
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        Cm = np.atleast_2d(np.asarray(C_, dtype=np.float64))
        if Cm.shape[0] != Cm.shape[1]:                                  # gaussian.py:35-36
            raise ParameterError("C has non-square shape {}".format(Cm.shape))
        if Cm.shape[1] != mu.shape[0]:                                  # gaussian.py:37-39
            raise ParameterError("mu and C have incompatible shapes {}, {}"
                                 .format(mu.shape, Cm.shape))
        self.mu = mu
        self.C = Cm
        try:
            self.L = np.linalg.cholesky(Cm)                             # gaussian.py:42
        except np.linalg.LinAlgError as e:
            raise ParameterError("C is not positive definite: {}".format(e))
        self.logdetC = 2 * np.sum(np.log(np.diag(self.L)))             # gaussian.py:43
        self.Ndim = len(mu)
        d = self.Ndim
        Linv = np.linalg.solve(self.L, np.eye(d))
        self.P = Linv.T @ Linv
        self.P = 0.5 * (self.P + self.P.T)
        self._Linv = np.ascontiguousarray(np.tril(Linv))
        h = C.c_void_p()
        Pc = np.ascontiguousarray(self.P)
        _lib.require_cuda()
        _lib.check(_lib.load().rmn_model_gaussian_create(
            C.byref(h), d, _lib.ptr(mu), _lib.ptr(Pc),
            _lib.ptr(self._Linv) if d <= _lib.SMALL_D_MAX else None, float(self.logdetC)))
        self._handle = h

    # the reference exposes the gradient under this name (gaussian.py:54-58); the
    # prior is flat so it is also the posterior gradient
    def grad_log_likelihood(self, theta):
        return self.grad_log_posterior(theta)

    def draw(self, Ndraws=1):
        """gaussian.py:60-66 (host-side convenience, not on the hot path)."""
        eps = np.random.normal(size=(Ndraws, self.mu.shape[0]))
        return np.dot(eps, self.L.T) + self.mu
