"""
riemann_b200 -- B200-native many-chain Metropolis-Hastings behind riemann's
Sampler / Proposal / Model API (reference: rscalzo/riemann, riemann/__init__.py:10-13).

Python here is host-side glue only: every MH iteration (model log-posterior and
gradient, proposal, Philox RNG, accept/reject, state swap) runs in hand-written CUDA
for sm_100a behind the C ABI of include/riemann_b200.h.  No CPU fallback.
"""
from .sampling_errors import ParameterError, RiemannBaseError
from .models.model import Model, DeviceModel, grad
from .proposals.proposal import Proposal, DeviceProposal
from .samplers.sampler import Sampler
from .samplers.ptsampler import PTSampler, TemperedModel

__all__ = ["Model", "Sampler", "Proposal", "ParameterError", "RiemannBaseError",
           "DeviceModel", "DeviceProposal", "PTSampler", "TemperedModel", "grad"]
