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

    def propose(self, theta):
        raise NotImplementedError(
            "device proposals are evaluated inside the fused MH kernels; drive them with "
            "riemann_b200.Sampler (Sampler.run_injected replays a given noise stream)")

    def __del__(self):
        try:
            if self._handle:
                _lib.load().rmn_proposal_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
