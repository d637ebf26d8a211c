import numpy as np, torch, sys, time
sys.path.insert(0, "/root/repo")
from riemann_b200 import Sampler
from riemann_b200.models import benchmarks
from riemann_b200.proposals.hamiltonian import MALA
for d, eps in ((100, 0.12), (1000, 0.08)):
    m = benchmarks.gauss_corr(d)
    K = 512
    rng = np.random.default_rng(0)
    th0 = rng.standard_normal((K, d)) * np.sqrt(0.1) + rng.standard_normal((K, 1)) * np.sqrt(0.9)
    s = Sampler(m, MALA(eps, m.grad_log_likelihood), th0, seed=3, precision="tf32x3")
    tot = 0
    for T in (1, 10, 100, 400, 600):
        s.run(T, trace=False); tot += T
        th, lp = s.state_tensors()
        want = m.log_posterior_batch(th).cpu().numpy()
        err = np.abs(lp.cpu().numpy() - want)
        print("d=%d after %4d steps: carried-lp max err %.3e  mean %.3e  acc %.3f" % (d, tot, err.max(), err.mean(), s.diagnostics(allreduce=False)["accept_rate"]))
    # per-step delta accuracy: injected, compare lp difference with fp64 at the device's own points
    s2 = Sampler(m, MALA(eps, m.grad_log_likelihood), th0[:64], seed=3, precision="tf32x3")
    T = 20
    ex = s2.run_injected(xi=rng.standard_normal((T, 64, d)), u=rng.uniform(size=(T, 64)))
    chain = s2._chain_thetas            # [T+1, 64, d]
    lpc = s2._chain_logpost
    cur64 = np.stack([m.log_posterior_batch(chain[t]).cpu().numpy() for t in range(T)])
    prop64 = np.stack([m.log_posterior_batch(ex["prop_theta"][t]).cpu().numpy() for t in range(T)])
    dd = (ex["prop_logpost"] - lpc[:T]) - (prop64 - cur64)
    print("d=%d per-step |delta-lp error| max %.3e mean %.3e" % (d, np.abs(dd).max(), np.abs(dd).mean()))
