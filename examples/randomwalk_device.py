#!/usr/bin/env python
"""
Device version of the reference's examples/test_randomwalk.py: the same benchmark target, the same proposal it
runs (AdaptScaleCovHMC, 5 leapfrog steps), the same chain length and burn-in, the same quantities printed --
acceptance fraction, adapted step-size scale, integrated autocorrelation time per parameter -- with the imports
pointing at riemann_b200 and emcee's estimator coming from riemann_b200.diagnostics.  A second block runs the
same target with 65,536 chains at once, which is what the engine is for.  Needs a B200 (no CPU fallback).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riemann_b200 import Sampler, diagnostics                       # noqa: E402
from riemann_b200.models.benchmarks import benchmark_gauss2d_corr   # noqa: E402
from riemann_b200.proposals.hamiltonian import AdaptScaleCovHMC     # noqa: E402
from riemann_b200.proposals.randomwalk import AdaptScaleRandomWalk  # noqa: E402


def single_chain():
    model = benchmark_gauss2d_corr
    proposal = AdaptScaleCovHMC(0.1, 5, model.grad_log_likelihood, np.eye(2), t_adapt=100, smooth_adapt=True)
    proposal.scale = 1.0
    sampler = Sampler(model, proposal, np.ones(2))
    sampler.run(10000, 1000, 1)
    chain = np.array(sampler._chain_thetas)
    tau = diagnostics.integrated_time(chain)
    accept = np.mean(np.any(chain[:-1] != chain[1:], axis=1))
    print("one chain: %d samples kept, acceptance fraction %.3f, scale %.3f" % (len(chain), accept, proposal.scale))
    print("           autocorrelation time per parameter:", np.round(tau, 1))
    print("           sample mean", np.round(chain.mean(axis=0), 3), "sample covariance", np.round(np.cov(chain.T), 3).tolist())


def many_chains(K=65536):
    model = benchmark_gauss2d_corr
    sampler = Sampler(model, AdaptScaleRandomWalk(1e-4 * np.eye(2)), np.ones(2), K=K, seed=1)
    sampler.run(2000, trace=False)               # burn-in: every chain adapts its own step-size scale
    sampler.reset_diagnostics()
    sampler.run(10000, trace=False)
    d = sampler.diagnostics()
    print("%d chains x %d steps: acceptance %.3f, max R-hat %.4f, min ESS %.3g, tau (steps) %s"
          % (d["chains"], d["steps"], d["accept_rate"], float(np.max(d["rhat"])), d["min_ess"], np.round(d["tau"], 1)))


if __name__ == "__main__":
    single_chain()
    many_chains()
