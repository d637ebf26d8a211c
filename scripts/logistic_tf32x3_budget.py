#!/usr/bin/env python
"""Measured accuracy budget of precision="tf32x3" on the logistic model at BASELINE config 4 / 5 shapes:
max and rms |log-posterior(tf32x3 sweep) - log-posterior(fp64 kernels)| at the same points, and the error of
the DIFFERENCE between a state and a MALA proposal from it (what enters the accept test)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riemann_b200 import synthetic
from riemann_b200 import Sampler
from riemann_b200.models.logistic import LogisticRegression
from riemann_b200.proposals.hamiltonian import MALA

res = {}
for name, N, d, K, eps in (("config4", 1000000, 100, 512, 0.02), ("config5", 100000, 64, 512, 0.05)):
    X, y, ts, pv = synthetic.logistic_problem(N, d)
    dm = LogisticRegression(X, y, pv)
    rng = np.random.default_rng(8)
    th0 = ts[None] + 0.01 * rng.standard_normal((K, d))
    xi, u = rng.standard_normal((1, K, d)), np.ones((1, K))
    s = Sampler(dm, MALA(eps, dm.grad_log_posterior), th0, precision="tf32x3")
    lp0 = np.asarray(s._chain_logpost[0]).copy()
    ex = s.run_injected(xi=xi, u=u)
    lp1, th1 = ex["prop_logpost"][0], ex["prop_theta"][0]
    w0 = dm.log_posterior_batch(th0).cpu().numpy()
    w1 = dm.log_posterior_batch(th1).cpu().numpy()
    e0, e1 = lp0 - w0, lp1 - w1
    res[name] = {"N": N, "d": d, "points": 2 * K, "logpost_magnitude": float(np.abs(w0).mean()),
                 "max_abs_err": float(max(np.abs(e0).max(), np.abs(e1).max())),
                 "rms_err": float(np.sqrt(np.mean(np.concatenate([e0, e1]) ** 2))),
                 "max_abs_err_of_difference": float(np.abs(e1 - e0).max()),
                 "mean_abs_difference": float(np.abs(w1 - w0).mean())}
print(json.dumps(res, indent=1))
