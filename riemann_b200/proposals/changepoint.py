"""ChangepointRegression1DProp -- examples/test_changepoint.py:18-73."""
import ctypes as C

import numpy as np

from .. import _lib
from ..models.changepoint import ChangepointRegression1D
from ..sampling_errors import ParameterError
from .proposal import DeviceProposal


class ChangepointRegression1DProp(DeviceProposal):
    """
    4-way mixture: RW on the changepoint locations / the step heights / the noise
    scale, or a birth/death move.  Block selection draws a FRESH uniform per elif
    (test_changepoint.py:48,51,54), so the effective probabilities are
    .20 / .32 / .288 / .192; the thresholds are kept as attributes like the reference.
    """

    _adaptive = False

    def __init__(self, model, hscale):
        if not isinstance(model, ChangepointRegression1D):
            raise ParameterError("ChangepointRegression1DProp needs a ChangepointRegression1D model")
        self.model = model
        self._model = model
        self.Ndata = len(model.x)
        self.hscale = hscale
        self.P_cumprop_cpx = 0.20
        self.P_cumprop_cpv = 0.40
        self.P_cumprop_sig = 0.60
        self.P_cumprop_dim = 1.00
        self.k = None

    def propose(self, theta):
        """
        examples/test_changepoint.py:44-73 for ONE state, evaluated by the device kernel.  The
        draws come from numpy's global stream in the reference's order (the host only routes
        them into the kernel's tape slots; no proposal arithmetic happens here).
        """
        from .. import _lib
        from ..models.changepoint import unpack_state
        from ..samplers.sampler import Sampler
        S = _lib.CP_SLOT
        k = len(theta.cpx)
        row = np.zeros(_lib.CP_NSLOT)
        row[S["acc"]] = 1.0                                  # log(1) = 0 is never < mhratio
        nnorm = None
        row[S["sel1"]] = np.random.uniform()
        if row[S["sel1"]] < self.P_cumprop_cpx:
            nnorm = k
        else:
            row[S["sel2"]] = np.random.uniform()
            if row[S["sel2"]] < self.P_cumprop_cpv:
                nnorm = k + 1
            else:
                row[S["sel3"]] = np.random.uniform()
                if row[S["sel3"]] < self.P_cumprop_sig:
                    nnorm = 1
                else:
                    birth = True
                    if k > 0:
                        row[S["bd"]] = np.random.uniform()
                        birth = row[S["bd"]] > 0.5
                    if birth:
                        row[S["s"]] = np.random.uniform(self.model.xmin, self.model.xmax)
                        row[S["du"]] = np.random.uniform(-0.1, 0.1)
                    else:
                        row[S["n"]] = np.random.randint(k)
        if nnorm:
            row[S["xi"]:S["xi"] + nnorm] = np.random.normal(size=(nnorm,))
        s = Sampler(self.model, self, theta)
        out = s.run(1, trace=False, inject={"tape": row[None, None, :]}, extras=True)
        th = unpack_state(out["prop_k"][0, 0], out["prop_cpx"][0, 0], out["prop_cpv"][0, 0],
                          out["prop_sig"][0, 0])
        return th, float(out["logqratio"][0, 0])

    def _create_handle(self, d):
        h = C.c_void_p()
        p = np.array([self.P_cumprop_cpx, self.P_cumprop_cpv, self.P_cumprop_sig], dtype=np.float64)
        _lib.check(_lib.load().rmn_proposal_changepoint_create(C.byref(h), float(self.hscale),
                                                               _lib.ptr(p)))
        return h
