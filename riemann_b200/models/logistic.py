"""
Bayesian logistic regression with a N(0, prior_var I) prior.  Not in the reference
(SURVEY.md row A11); follows the Model protocol (riemann/models/model.py:27-55).
"""
import ctypes as C

import numpy as np

from .. import _lib
from ..sampling_errors import ParameterError
from .model import DeviceModel


class LogisticRegression(DeviceModel):
    def __init__(self, X, y, prior_var=100.0):
        DeviceModel.__init__(self)
        torch = _lib.require_cuda()
        Xt = X if isinstance(X, torch.Tensor) else torch.as_tensor(np.asarray(X))
        yt = y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y))
        if Xt.dim() != 2 or yt.dim() != 1 or Xt.shape[0] != yt.shape[0]:
            raise ParameterError("X and y have incompatible shapes {}, {}"
                                 .format(tuple(Xt.shape), tuple(yt.shape)))
        if not prior_var > 0:
            raise ParameterError("prior_var must be positive")
        self.X = Xt.to(device="cuda", dtype=torch.float64).contiguous()
        self.y = yt.to(device="cuda", dtype=torch.float64).contiguous()
        self.prior_var = float(prior_var)
        self.N, self.Ndim = int(self.X.shape[0]), int(self.X.shape[1])
        h = C.c_void_p()
        _lib.check(_lib.load().rmn_model_logistic_create(
            C.byref(h), self.N, self.Ndim, _lib.ptr(self.X), _lib.ptr(self.y), self.prior_var))
        self._handle = h

    def metric_batch(self, Theta):
        torch, t = self._as_device(Theta)
        out = torch.empty((t.shape[0], self.Ndim, self.Ndim), dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().rmn_model_metric(self._handle, t.shape[0], _lib.ptr(t), _lib.ptr(out),
                                                _lib.stream_ptr()))
        return out

    def metric(self, theta):
        return self.metric_batch(np.atleast_1d(theta))[0].cpu().numpy()
