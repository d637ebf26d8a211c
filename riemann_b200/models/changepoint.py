"""ChangepointRegression1D / ChangepointParams -- riemann/models/changepoint.py:22-240."""
import ctypes as C

import numpy as np

from .. import _lib
from ..sampling_errors import ParameterError
from .model import DeviceModel

LANES = _lib.CP_LANES


class ChangepointParams(object):
    """State (cpx[k], cpv[k+1], sig) of the model (changepoint.py:22-45)."""

    def __init__(self, cpx, cpv, sig):
        if len(cpv) != len(cpx) + 1:                                    # changepoint.py:35-39
            raise ValueError("number of constant pieces ({}) must be "
                             "1 more than number of changepoints ({})"
                             .format(len(cpv), len(cpx)))
        self.cpx = np.array(cpx, dtype=np.float64)
        self.cpv = np.array(cpv, dtype=np.float64)
        self.sig = np.array(sig, dtype=np.float64)

    def __str__(self):
        return ("ChangepointParams instance with sig = {}, "
                "cpx = {}, cpv = {}".format(self.sig, self.cpx, self.cpv))


def pack_states(thetas):
    """list of ChangepointParams -> canonical padded arrays (k, cpx, cpv, sig)."""
    n = len(thetas)
    k = np.zeros(n, dtype=np.int32)
    cpx = np.zeros((n, LANES))
    cpv = np.zeros((n, LANES))
    sig = np.zeros(n)
    for i, th in enumerate(thetas):
        ki = len(th.cpx)
        if ki > LANES - 1:
            raise ParameterError("at most {} changepoints per chain are supported (got {})"
                                 .format(LANES - 1, ki))
        k[i] = ki
        cpx[i, :ki] = th.cpx
        cpv[i, :ki + 1] = th.cpv
        sig[i] = float(np.squeeze(th.sig))
    return k, cpx, cpv, sig


def unpack_state(k, cpx, cpv, sig):
    k = int(k)
    return ChangepointParams(cpx[:k].copy(), cpv[:k + 1].copy(), float(sig))


class ChangepointRegression1D(DeviceModel):
    """
    1-D piecewise-constant regression (changepoint.py:81-181).  x must be sorted
    ascending (the reference's own data generator sorts it, :170); the engine exploits
    that to evaluate the likelihood from prefix sums in O(k log M).
    """

    def __init__(self, x, y, xmin, xmax, lamb, kmax, alpha, beta, debug=False):
        DeviceModel.__init__(self)
        if len(x) != len(y):                                            # changepoint.py:91-95
            raise ValueError("length of predictor array ({}) must be"
                             "same as length of response array ({})".format(len(x), len(y)))
        self.xmin, self.xmax = xmin, xmax
        self.kmax = kmax                     # stored, never enforced (changepoint.py:100)
        self.lamb, self.alpha, self.beta = lamb, alpha, beta
        self.debug = debug
        self.Ndim = 2 * LANES + 1
        self._x = np.array(x, dtype=np.float64)
        self._y = np.array(y, dtype=np.float64)
        if len(self._x):
            self._upload()

    def _upload(self):
        if np.any(np.diff(self._x) < 0):
            raise ParameterError("ChangepointRegression1D: x must be sorted ascending")
        if self._handle:
            _lib.load().rmn_model_destroy(self._handle)
            self._handle = None
        _lib.require_cuda()
        h = C.c_void_p()
        x = np.ascontiguousarray(self._x)
        y = np.ascontiguousarray(self._y)
        _lib.check(_lib.load().rmn_model_changepoint_create(
            C.byref(h), len(x), _lib.ptr(x), _lib.ptr(y), float(self.xmin), float(self.xmax),
            float(self.lamb), int(self.kmax), float(self.alpha), float(self.beta)))
        self._handle = h

    # the reference's example assigns model.x, model.y after construction
    # (examples/test_changepoint.py:146); re-upload when both are in place
    @property
    def x(self):
        return self._x

    @x.setter
    def x(self, v):
        self._x = np.array(v, dtype=np.float64)
        if len(self._x) and len(self._x) == len(self._y):
            self._upload()

    @property
    def y(self):
        return self._y

    @y.setter
    def y(self, v):
        self._y = np.array(v, dtype=np.float64)
        if len(self._y) and len(self._x) == len(self._y):
            self._upload()

    def _eval_states(self, thetas, which):
        torch = _lib.require_cuda()
        if self._handle is None:
            raise ParameterError("ChangepointRegression1D has no data yet")
        k, cpx, cpv, sig = pack_states(thetas)
        dk, dx, dv, ds = (torch.as_tensor(a, device="cuda") for a in (k, cpx, cpv, sig))
        out = torch.empty(len(thetas), dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().rmn_model_cp_logpost(
            self._handle, which, len(thetas), _lib.ptr(dk), _lib.ptr(dx), _lib.ptr(dv),
            _lib.ptr(ds), _lib.ptr(out), _lib.stream_ptr()))
        return out.cpu().numpy()

    def log_posterior_batch(self, thetas):
        return self._eval_states(thetas, 0)

    def log_likelihood(self, theta):
        return float(self._eval_states([theta], 1)[0])

    def log_prior(self, theta):
        return float(self._eval_states([theta], 2)[0])

    def log_posterior(self, theta):
        return float(self._eval_states([theta], 0)[0])

    def predict(self, theta, x):
        """changepoint.py:175-181 (host-side helper for posterior-predictive plots)."""
        return theta.cpv[np.searchsorted(theta.cpx, x)]

    def generate_synthetic_data(self, theta, Ndata):
        """changepoint.py:162-173 (host-side, numpy global stream like the reference)."""
        L = self.xmax - self.xmin
        x = np.sort(self.xmin + L * np.random.uniform(size=(Ndata,)))
        epsilon = np.random.normal(size=x.shape)
        return x, self.predict(theta, x) + theta.sig * epsilon
