#!/usr/bin/env python
"""GPU check of the fused tcgen05 likelihood sweep (precision="tf32x3", logistic_fused.cu) against the fp64 kernels at
the same points: log-posterior error (max / rms / mean = bias), error of the DIFFERENCE state -> MALA proposal (what enters
the accept test) and the relative error of the gradient recovered from the proposal's drift.
    python scripts/lg_fused_check.py small|config5|config4 [...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riemann_b200 import Sampler, synthetic                      # noqa: E402
from riemann_b200.models.logistic import LogisticRegression      # noqa: E402
from riemann_b200.proposals.hamiltonian import MALA               # noqa: E402

SHAPES = {"tiny": (4096, 16, 200, 0.05), "ragged": (10037, 100, 131, 0.05), "d64": (20000, 64, 256, 0.05),
          "config5": (100000, 64, 512, 0.05), "config4": (1000000, 100, 512, 0.02)}
res = {}
for name in sys.argv[1:] or ["tiny"]:
    N, d, K, eps = SHAPES[name]
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
    g0 = dm.grad_log_posterior_batch(th0).cpu().numpy()
    g_dev = (th1 - th0 - eps * xi[0]) * 2.0 / (eps * eps)        # theta' = theta + eps xi + eps^2/2 g  (hamiltonian.py:27-40)
    e0, e1 = lp0 - w0, lp1 - w1
    res[name] = {"N": N, "d": d, "points": 2 * K, "logpost_magnitude": float(np.abs(w0).mean()),
                 "max_abs_err": float(max(np.abs(e0).max(), np.abs(e1).max())),
                 "rms_err": float(np.sqrt(np.mean(np.concatenate([e0, e1]) ** 2))),
                 "mean_err": float(np.mean(np.concatenate([e0, e1]))),
                 "max_abs_err_of_difference": float(np.abs(e1 - e0).max()),
                 "mean_abs_difference": float(np.abs(w1 - w0).mean()),
                 "grad_rel_err": float(np.max(np.abs(g_dev - g0)) / np.max(np.abs(g0))),
                 "grad_scale": float(np.max(np.abs(g0)))}
    print(name, json.dumps(res[name]), flush=True)
