"""
Stated accuracy budgets of the tensor-core precision modes -- ONE number per mode and shape, quoted by
include/riemann_b200.h and DESIGN.md and asserted by the tests (tests/test_gpu_logistic.py,
tests/test_gpu_dense_tf32.py).  Everything here is about what enters the accept test of
riemann/samplers/sampler.py:83-84: the log-posterior DIFFERENCE proposal - state.

f64 (default) has no budget: it is the reference's arithmetic (1e-9 relative against the numpy oracle).
"""
import math


def logistic_tf32x3_difference(N):
    """|(logpost' - logpost)_tf32x3 - (logpost' - logpost)_f64| for a state and a proposal from it: 1e-3 at
    BASELINE config 4 (N = 1e6) -- SURVEY 8d's gate -- scaling with sqrt(N) (independent rounding errors of the
    fp32-accurate logits), never below 5e-5.  Measured with the scaled-fp16 correction terms of the fused sweep:
    8.3e-4 / 1.9e-4 / 6.2e-5 at N = 1e6 / 1e5 / 2e4 (TF32 correction terms: 1.2e-3 / 2.7e-4 / 7e-5; bf16 ones:
    1.7e-3 / 4.0e-4 / 1.9e-4)."""
    return max(5e-5, 1e-3 * math.sqrt(N / 1.0e6))


def logistic_tf32x3_offset(N):
    """|logpost_tf32x3 - logpost_f64| of ONE state.  The fp32 softplus of the fused sweep carries a constant
    bias of ~1.5e-8 per data row (measured 0.015 at N = 1e6); it is the same for every state, so it cancels in
    every Metropolis-Hastings ratio, and is bounded here only to catch real errors: 5e-8 per row."""
    return 5e-8 * N + 1e-4


def dense_tf32x3_difference(d):
    """Dense Gaussian model, precision="tf32x3": |(logpost' - logpost)_device - (logpost' - logpost)_f64| at the device's
    own points: 5e-4 at BASELINE config 3 (d = 1000; measured 3.3e-4), 5e-5 at d = 100 (measured 1.6e-5)."""
    return 5e-7 * d


def dense_tf32x3_carried(d):
    """The CARRIED log-posterior against a fresh fp64 evaluation between two exact refreshes (every 512 steps):
    2e-2 at d = 1000 (measured 9e-3, a stable offset: the precision matrix rounded to fp32 is the model), 2e-3 at
    d = 100.  Like the logistic offset it is common to state and proposal and cancels in the accept test."""
    return 2e-5 * d
