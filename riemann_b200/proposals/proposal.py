"""
Proposal protocol -- riemann/proposals/proposal.py:1-26.

In the engine a proposal is a set of hyper-parameters; ``propose`` runs INSIDE the
fused device kernels (together with the model evaluation and the accept/reject), for
all chains at once.  ``DeviceProposal`` subclasses therefore carry only the numbers
the kernel needs and create a C handle on demand.
"""
from .. import _lib


class Proposal(object):
    def __init__(self):
        pass

    def propose(self, theta):
        """
        :return theta_p:  proposed parameter vector  ~ q(theta'|theta)
        :return logqratio:  log(q(theta'|theta)/q(theta|theta'))   (subtracted by the Sampler)
        """
        raise NotImplementedError("Non-overloaded abstract method!")

    def adapt(self, theta):
        pass


class DeviceProposal(Proposal):
    """A proposal with a device kernel."""

    _handle = None
    _model = None          # device model the proposal needs (gradient / metric), if any

    def _create_handle(self, d):
        raise NotImplementedError

    def _get_handle(self, d):
        if self._handle is None:
            self._handle = self._create_handle(d)
        return self._handle

    def _propose_model(self, d):
        """Model the one-step device sampler runs on: the proposal's own (gradient / metric)
        model, or -- for model-free proposals (RW, pCN) -- a standard normal placeholder whose
        value never enters the proposal."""
        if self._model is not None:
            return self._model
        if getattr(self, "_placeholder", None) is None or self._placeholder.Ndim != d:
            from ..models.gaussian import MultiGaussianDist
            import numpy as np
            self._placeholder = MultiGaussianDist(np.zeros(d), np.eye(d))
        return self._placeholder

    def propose(self, theta):
        """
        proposal.py:10-17 for ONE point, evaluated by the device kernels: draws xi from numpy's
        global stream exactly like the reference (randomwalk.py:25, hamiltonian.py:79), injects it
        into a one-step device run that cannot accept (u = 1), and returns
        (theta', log q(theta'|theta)/q(theta|theta')).  For the hot path use Sampler.run.
        """
        import numpy as np
        from ..samplers.sampler import Sampler
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        d = theta.shape[0]
        s = Sampler(self._propose_model(d), self, theta)
        if getattr(self, "_adaptive", False):
            s.set_adapt(np.mean(self.scale), 1, 1)          # use the current adapted scale
        xi = np.random.normal(size=theta.shape)
        keep = {k: getattr(self, k) for k in ("scale", "Nsamples", "Naccepts", "accept_rate", "eps")
                if hasattr(self, k)}
        out = s.run(1, trace=False, inject={"xi": xi[None, None, :], "u": np.ones((1, 1))}, extras=True)
        for k, v in keep.items():                            # a dry proposal must not adapt
            setattr(self, k, v)
        return out["prop_theta"][0, 0], float(out["logqratio"][0, 0])

    def __del__(self):
        try:
            if self._handle:
                _lib.load().rmn_proposal_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
