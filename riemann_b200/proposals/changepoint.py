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

    def _create_handle(self, d):
        h = C.c_void_p()
        p = np.array([self.P_cumprop_cpx, self.P_cumprop_cpv, self.P_cumprop_sig], dtype=np.float64)
        _lib.check(_lib.load().rmn_proposal_changepoint_create(C.byref(h), float(self.hscale),
                                                               _lib.ptr(p)))
        return h
