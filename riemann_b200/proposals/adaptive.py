"""AdaptScaleProposal mix-in -- riemann/proposals/adaptive.py:11-35.

The adaptation itself runs per chain inside the device kernels (common.cuh
AdaptState::update).  This class keeps the reference's public attributes
(scale, accept_rate, Nsamples, Naccepts, target_accept_rate); the Sampler refreshes
them from the device after every batch -- scalars for one chain, arrays for K chains.
"""
from .proposal import DeviceProposal


class AdaptScaleProposal(DeviceProposal):
    def __init__(self, target_accept_rate):
        self.Nsamples = 0
        self.Naccepts = 0
        self.last_theta = None
        self.accept_rate = 0.0
        self.target_accept_rate = target_accept_rate
        self.scale = 1.0
        self._adaptive = True
