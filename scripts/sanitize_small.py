"""One small run of every device sampler family on awkward shapes (chain counts and dimensions that
are not multiples of any tile, warp or vector width):

    python scripts/sanitize_small.py

Every kernel of the library launches at least once; each case must finish without a CUDA fault and
with finite log-posteriors.  Numbers are not checked here -- the parity tests do that.  (It was
written to run under `compute-sanitizer --tool memcheck`; that tool is closed on the GPU pool this
repo is developed on, so out-of-bounds accesses are hunted with ragged-shape parity tests instead:
tests/test_gpu_*.py `ragged` cases.)  Combinations the library refuses print `refused:` and the
message."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    torch.cuda.set_device(0)
    from riemann_b200 import Sampler, PTSampler
    from riemann_b200.models.gaussian import MultiGaussianDist
    from riemann_b200.models.logistic import LogisticRegression
    from riemann_b200.models.changepoint import ChangepointParams, ChangepointRegression1D
    from riemann_b200.proposals.changepoint import ChangepointRegression1DProp
    from riemann_b200.proposals import randomwalk as rw, hamiltonian as hm
    from riemann_b200 import synthetic

    rng = np.random.default_rng(0)

    def spd(d, s=1.0):
        A = rng.standard_normal((d, d))
        return s * (A @ A.T / d + 0.3 * np.eye(d))

    def run(tag, make, T=6, **kw):
        try:
            s = make()
        except Exception as e:                      # a combination the library refuses (says so loudly) is not a memory error
            print("%-34s refused: %s" % (tag, str(e).splitlines()[0][:90]), flush=True)
            return
        s.run(T, **kw)
        dg = s.diagnostics(allreduce=False)
        lp = np.asarray(s._chain_logpost[-1] if kw.get("trace", True) else s._chain_logpost)
        assert np.all(np.isfinite(lp)), tag
        print("%-34s ok  accept=%.2f" % (tag, dg["accept_rate"]), flush=True)

    # changepoint, K not a multiple of the 8 chains a warp holds
    c = synthetic.changepoint_problem()
    cm = ChangepointRegression1D(c["x"], c["y"], c["xmin"], c["xmax"], c["lamb"], c["kmax"], c["alpha"], c["beta"])
    run("changepoint K=203", lambda: Sampler(cm, ChangepointRegression1DProp(cm, c["hscale"]),
                                             ChangepointParams(*c["theta0"]), K=203, seed=1), T=40, trace=False)

    # small-d Gaussian family
    for d in (1, 2, 5, 8):
        C = spd(d)
        g = MultiGaussianDist(rng.standard_normal(d), C)
        th0 = rng.standard_normal((77, d))
        run("small d=%d RW" % d, lambda: Sampler(g, rw.MetropolisRandomWalk(0.5 * C), th0, seed=2))
        run("small d=%d AdaptScaleCovRW" % d, lambda: Sampler(g, rw.AdaptScaleCovRandomWalk(0.5 * C), th0, seed=2), T=12)
        run("small d=%d HMC3 mass" % d, lambda: Sampler(g, hm.VanillaHMC(0.2, 3, g.grad_log_likelihood, M=np.linalg.inv(C)), th0, seed=2))
        run("small d=%d AdaptScaleCovHMC" % d, lambda: Sampler(g, hm.AdaptScaleCovHMC(0.2, 2, g.grad_log_likelihood, M0=np.eye(d)), th0, seed=2))
        gz = MultiGaussianDist(np.zeros(d), C)
        run("small d=%d AdaptScalepCN" % d, lambda: Sampler(gz, rw.AdaptScalepCN(C, 0.8), th0, seed=2))
    g2 = MultiGaussianDist(np.zeros(2), spd(2))
    pt = PTSampler(g2, rw.MetropolisRandomWalk(0.4 * np.eye(2)), np.ones(2), K=37, seed=4)
    pt.run(25)
    print("%-34s ok" % "PTSampler K=37", flush=True)

    # dense Gaussian family, d not a multiple of 16, K not a multiple of 128
    for d in (12, 37, 100):
        C = spd(d)
        g = MultiGaussianDist(rng.standard_normal(d), C)
        th0 = 0.3 * rng.standard_normal((131, d))
        run("dense d=%d RW diag" % d, lambda: Sampler(g, rw.MetropolisRandomWalk(0.1 * np.eye(d)), th0, seed=3))
        run("dense d=%d AdaptScaleRW dense" % d, lambda: Sampler(g, rw.AdaptScaleRandomWalk(0.1 * C), th0, seed=3))
        run("dense d=%d MALA" % d, lambda: Sampler(g, hm.MALA(0.1, g.grad_log_likelihood), th0, seed=3))
        run("dense d=%d HMC3" % d, lambda: Sampler(g, hm.AdaptScaleHMC(0.1, 3, g.grad_log_likelihood), th0, seed=3))
        run("dense d=%d HMC3 mass" % d, lambda: Sampler(g, hm.VanillaHMC(0.2, 3, g.grad_log_likelihood, M=np.linalg.inv(C)), th0, seed=3))
        gz = MultiGaussianDist(np.zeros(d), C)
        run("dense d=%d pCN" % d, lambda: Sampler(gz, rw.pCN(C, 0.9), th0, seed=3))
        run("dense d=%d MALA tf32x3" % d, lambda: Sampler(g, hm.MALA(0.1, g.grad_log_likelihood), th0, seed=3, precision="tf32x3"))
        run("dense d=%d RW tf32x3" % d, lambda: Sampler(g, rw.MetropolisRandomWalk(0.1 * np.eye(d)), th0, seed=3, precision="tf32x3"))

    # logistic family, N and d ragged
    for (N, d) in ((333, 7), (1500, 20)):
        X, y, ts, pv = synthetic.logistic_problem(N, d, seed=5)
        lm = LogisticRegression(X, y, pv)
        th0 = ts[None] + 0.05 * rng.standard_normal((45, d))
        for prec in ("f64", "tf32x3"):
            run("logistic N=%d d=%d MALA %s" % (N, d, prec), lambda: Sampler(lm, hm.MALA(0.05, lm.grad_log_posterior), th0, seed=6, precision=prec))
            run("logistic N=%d d=%d mMALA %s" % (N, d, prec), lambda: Sampler(lm, hm.SimplifiedMMALA(0.5, lm), th0, seed=6, precision=prec))
        run("logistic N=%d d=%d HMC3" % (N, d), lambda: Sampler(lm, hm.VanillaHMC(0.03, 3, lm.grad_log_posterior), th0, seed=6))
        run("logistic N=%d d=%d mMALA tf32-metric" % (N, d), lambda: Sampler(lm, hm.SimplifiedMMALA(0.5, lm), th0, seed=6, precision="tf32-metric"))
        run("logistic N=%d d=%d RW" % (N, d), lambda: Sampler(lm, rw.MetropolisRandomWalk(0.001 * np.eye(d)), th0, seed=6))
        run("logistic N=%d d=%d pCN tf32x3" % (N, d), lambda: Sampler(lm, rw.pCN(pv * np.eye(d), 0.995), th0, seed=6, precision="tf32x3"))
        M = spd(d, 50.0)
        run("logistic N=%d d=%d HMC3 mass" % (N, d), lambda: Sampler(lm, hm.VanillaHMC(0.2, 3, lm.grad_log_posterior, M=M), th0, seed=6))
        run("logistic N=%d d=%d AdaptMALA mass tf32x3" % (N, d), lambda: Sampler(lm, hm.AdaptScaleHMC(0.2, 1, lm.grad_log_posterior, M=M), th0, seed=6, precision="tf32x3"))
    print("sanitize_small: all cases ran")


if __name__ == "__main__":
    main()
