"""GPU check of the shared move schedule of the changepoint kernel (rmn_sampler_set_move_schedule): with 65,536 chains,
are the ergodic averages of the 8 chains of a schedule group correlated?  For each tracked functional the intra-group
correlation rho of the per-chain means is estimated from  Var(group mean) = Var(chain mean) (1 + 7 rho) / 8  over the
8,192 groups (standard error sqrt(2 / (8 * 7 * G)) = 0.0021 under independence) -- for the shared ("group") schedule
and, as the control, for per-chain schedules.  Also prints the pooled posterior means under both schedules.

    python scripts/cp_schedule_check.py [K] [window_steps] > gpurun_out/cp_schedule_check.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (build_workload)


def run(schedule, K, window, seed):
    os.environ["RMN_CP_SCHEDULE"] = schedule
    s, _ = bench.build_workload("changepoint", K, seed, 0, "f64")
    s.run(60000, trace=False)
    s.reset_diagnostics()
    for _ in range(window // 1000):
        s.run(1000, trace=False)
    m, v = s.chain_moments()                       # [8, K]
    G = K // 8
    out = {"schedule": schedule, "chains": K, "window_steps": window, "se_rho": float(np.sqrt(2.0 / (8 * 7 * G)))}
    rho, pair = [], []
    for f in range(m.shape[0]):
        x = m[f]
        vc = x.var(ddof=1)
        vg = x.reshape(G, 8).mean(axis=1).var(ddof=1)
        rho.append(float((8.0 * vg / vc - 1.0) / 7.0))
        a = x.reshape(G, 8)
        pair.append(float(np.corrcoef(a[:, 0], a[:, 1])[0, 1]))
    out["functionals"] = bench.FUNC_NAMES["changepoint"]
    out["intra_group_rho"] = rho
    out["z"] = [r / out["se_rho"] for r in rho]
    out["neighbour_pair_corr"] = pair
    out["pooled_mean"] = [float(t) for t in m.mean(axis=1)]
    out["pooled_mean_se"] = [float(t) for t in m.std(axis=1, ddof=1) / np.sqrt(K)]
    d = s.diagnostics(allreduce=False)
    out["accept_rate"] = float(d["accept_rate"])
    return out


if __name__ == "__main__":
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    res = [run("group", K, W, 20261018), run("chain", K, W, 20261018), run("group", K, W, 777)]
    print(json.dumps(res, indent=1))
