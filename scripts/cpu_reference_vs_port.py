"""Build container only (needs /root/reference): single-core speed of the UNMODIFIED reference sampler against the
numpy port that bench.py times as `cpu_baseline` (kind "port" -- the reference cannot travel to the GPU box), on the
headline changepoint workload.  Calibrates the CPU arm: how much faster or slower the port is than the real thing.

    python scripts/cpu_reference_vs_port.py [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, riemann_port as port      # noqa: E402   (a measurement of the CPU arm itself)


def rate(sampler, seconds, chunk=500):
    n, t0 = 0, time.perf_counter()
    with np.errstate(all="ignore"):
        while time.perf_counter() - t0 < seconds:
            for _ in range(chunk):
                sampler.sample()
            n += chunk
    return n / (time.perf_counter() - t0)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    if not refshim.reference_available():
        sys.exit("needs the reference tree")
    R = refshim.load_reference()
    pm, pprop, pth0, _ = port.make_changepoint_problem()
    np.random.seed(1)
    model = R.ChangepointRegression1D(pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb, pm.kmax, pm.alpha, pm.beta)
    prop = R.ChangepointRegression1DProp(model, pprop.hscale)
    theta0 = R.ChangepointParams(np.array(pth0.cpx), np.array(pth0.cpv), pth0.sig)
    with refshim.quiet():
        r_ref = rate(R.Sampler(model, prop, theta0), seconds)
    np.random.seed(1)
    r_port = rate(port.Sampler(pm, pprop, pth0), seconds)
    print("changepoint, one core, %.0f s each: reference %.0f steps/s, port %.0f steps/s, port/reference = %.2f"
          % (seconds, r_ref, r_port, r_port / r_ref))
    # BASELINE config 0/1: benchmark_gauss2d_corr with a random walk
    np.random.seed(1)
    g = R.benchmarks.benchmark_gauss2d_corr
    r_ref = rate(R.Sampler(g, R.MetropolisRandomWalk(0.5 * np.eye(2)), np.ones(2)), seconds)
    np.random.seed(1)
    r_port = rate(port.Sampler(port.benchmark_gauss(2), port.MetropolisRandomWalk(0.5 * np.eye(2)), np.ones(2)), seconds)
    print("gauss2d random walk, one core, %.0f s each: reference %.0f steps/s, port %.0f steps/s, port/reference = %.2f"
          % (seconds, r_ref, r_port, r_port / r_ref))


if __name__ == "__main__":
    main()
