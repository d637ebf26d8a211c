"""
TEST INFRASTRUCTURE -- build-container only.

Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/refshim.py) under np.random.seed(...)
while recording every draw it takes from numpy's global stream.  Each fixture
holds the injected stream (normals + uniforms) and the chain the reference
produced from it; the numpy port (oracle/riemann_port.py) and the CUDA engine
both replay the stream and must reproduce the chain.

    python -m oracle.gen_golden            # rewrites every fixture

The fixtures are committed; /root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np

from oracle import refshim
from oracle import riemann_port as port

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")
LANES = 16


def _run_recorded(R, sampler, T):
    """Run T reference steps; return per-step draw logs + proposals."""
    props, prop_lp = [], []
    model = sampler.model
    orig = model.log_posterior

    def rec_lp(theta):
        v = orig(theta)
        props.append(theta)
        prop_lp.append(v)
        return v

    model.log_posterior = rec_lp           # instance attribute; reference code untouched
    steps = []
    try:
        with refshim.quiet(), refshim.RecordingRNG() as rec:
            for _ in range(T):
                a = rec.mark()
                sampler.sample()
                steps.append(rec.log[a:rec.mark()])
    finally:
        del model.log_posterior
    return steps, props, np.array(prop_lp, dtype=np.float64)


def _vector_fixture(R, name, model, proposal, theta0, T, seed, extra=None, track_scale=False):
    np.random.seed(seed)
    sampler = R.Sampler(model, proposal, theta0)
    d = len(np.atleast_1d(theta0))
    scales = [getattr(proposal, "scale", 1.0)]
    if track_scale:
        orig_adapt = proposal.adapt

        def adapt(theta):
            orig_adapt(theta)
            scales.append(proposal.scale)
        proposal.adapt = adapt
    steps, props, prop_lp = _run_recorded(R, sampler, T)
    xi = np.zeros((T, d))
    u = np.zeros(T)
    for t, log in enumerate(steps):
        assert [k for k, _ in log] == ["normal", "uniform"], log
        xi[t] = log[0][1]
        u[t] = float(log[1][1])
    out = dict(xi=xi, u=u,
               thetas=np.array([np.atleast_1d(th) for th in sampler._chain_thetas], dtype=np.float64),
               logpost=np.array(sampler._chain_logpost, dtype=np.float64),
               prop_thetas=np.array([np.atleast_1d(p) for p in props], dtype=np.float64),
               prop_logpost=prop_lp, seed=np.int64(seed))
    if track_scale:
        out["scales"] = np.array(scales, dtype=np.float64)
        out["accept_rate"] = np.float64(proposal.accept_rate)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    acc = np.mean(np.any(out["thetas"][1:] != out["thetas"][:-1], axis=1))
    print("%-28s T=%d d=%d accept=%.3f" % (name, T, d, acc))


def _cp_pack_state(theta, lanes=LANES):
    k = len(theta.cpx)
    cpx = np.zeros(lanes)
    cpv = np.zeros(lanes)
    cpx[:k] = theta.cpx
    cpv[:k + 1] = theta.cpv
    return k, cpx, cpv, float(np.squeeze(theta.sig))


def _cp_tape_row(log, k, lanes=LANES):
    """Map one step's recorded draws onto the fixed slots (control flow of
    examples/test_changepoint.py:44-73 + sampler.py:84)."""
    row = np.zeros(port.cp_nslot(lanes))
    it = iter(log)

    def nxt(kind):
        kk, v = next(it)
        assert kk == kind, (kk, kind, log)
        return v

    u1 = float(nxt("uniform"))
    row[port.CP_SLOT_SEL1] = u1
    nnorm = None
    if u1 < 0.20:
        nnorm = k
    else:
        u2 = float(nxt("uniform"))
        row[port.CP_SLOT_SEL2] = u2
        if u2 < 0.40:
            nnorm = k + 1
        else:
            u3 = float(nxt("uniform"))
            row[port.CP_SLOT_SEL3] = u3
            if u3 < 0.60:
                nnorm = 1
            else:
                birth = True
                if k > 0:
                    ub = float(nxt("uniform"))
                    row[port.CP_SLOT_BD] = ub
                    birth = ub > 0.5
                if birth:
                    row[port.CP_SLOT_S] = float(nxt("uniform"))
                    row[port.CP_SLOT_DU] = float(nxt("uniform"))
                else:
                    row[port.CP_SLOT_N] = float(nxt("randint"))
    if nnorm is not None:
        v = np.atleast_1d(nxt("normal"))
        assert len(v) == nnorm, (len(v), nnorm)
        row[port.CP_SLOT_XI:port.CP_SLOT_XI + nnorm] = v
    row[port.CP_SLOT_ACC] = float(nxt("uniform"))
    assert next(it, None) is None
    return row


def _changepoint_fixture(R, name, nchains, T, seed0):
    pm, pprop, ptheta0, ptrue = port.make_changepoint_problem()
    tapes, ks, cpxs, cpvs, sigs, lps, plps = [], [], [], [], [], [], []
    for c in range(nchains):
        model = R.ChangepointRegression1D(pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb,
                                          pm.kmax, pm.alpha, pm.beta)
        prop = R.ChangepointRegression1DProp(model, pprop.hscale)
        theta0 = R.ChangepointParams(ptheta0.cpx, ptheta0.cpv, ptheta0.sig)
        np.random.seed(seed0 + c)
        sampler = R.Sampler(model, prop, theta0)
        with np.errstate(all="ignore"):
            steps, props, prop_lp = _run_recorded(R, sampler, T)
        chain = sampler._chain_thetas
        assert len(chain) == T + 1
        tape = np.stack([_cp_tape_row(steps[t], len(chain[t].cpx)) for t in range(T)])
        st = [_cp_pack_state(th) for th in chain]
        assert max(s[0] for s in st) < LANES
        tapes.append(tape)
        ks.append([s[0] for s in st])
        cpxs.append([s[1] for s in st])
        cpvs.append([s[2] for s in st])
        sigs.append([s[3] for s in st])
        lps.append(sampler._chain_logpost)
        plps.append(prop_lp)
        k_arr = np.array(ks[-1])
        print("%-28s chain %d T=%d k in [%d,%d] accept=%.3f" % (
            name, c, T, k_arr.min(), k_arr.max(),
            np.mean(np.array(lps[-1])[1:] != np.array(lps[-1])[:-1])))
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name + ".npz"),
        tape=np.array(tapes), k=np.array(ks, dtype=np.int32), cpx=np.array(cpxs),
        cpv=np.array(cpvs), sig=np.array(sigs), logpost=np.array(lps, dtype=np.float64),
        prop_logpost=np.array(plps), x=pm.x, y=pm.y,
        hyper=np.array([pm.xmin, pm.xmax, pm.lamb, pm.kmax, pm.alpha, pm.beta, pprop.hscale]),
        seed0=np.int64(seed0))


def _cp_summary_worker(seed):
    """One long chain of the UNMODIFIED reference on the bench problem; window statistics."""
    R = refshim.load_reference()
    pm, pprop, ptheta0, _ = port.make_changepoint_problem()
    model = R.ChangepointRegression1D(pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb, pm.kmax, pm.alpha, pm.beta)
    prop = R.ChangepointRegression1DProp(model, pprop.hscale)
    np.random.seed(seed)
    s = R.Sampler(model, prop, R.ChangepointParams(ptheta0.cpx, ptheta0.cpv, ptheta0.sig))
    with refshim.quiet(), np.errstate(all="ignore"):
        s.run(10000, 6000)
    ks = np.array([len(t.cpx) for t in s._chain_thetas])
    sig = np.array([float(np.squeeze(t.sig)) for t in s._chain_thetas])
    lp = np.array(s._chain_logpost)
    return (ks.mean(), sig.mean(), np.mean(lp[1:] != lp[:-1]), np.bincount(ks, minlength=LANES)[:LANES] / len(ks))


def _changepoint_posterior_summary(name, nchains=48, seed0=7000):
    """Distributional fixture: per-chain window means (steps 6000..10000 from the example's
    start state) of 48 independent reference chains, each on numpy's own stream."""
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(_cp_summary_worker, [seed0 + i for i in range(nchains)])
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"),
                        mean_k=np.array([r[0] for r in res]), mean_sig=np.array([r[1] for r in res]),
                        accept=np.array([r[2] for r in res]), khist=np.array([r[3] for r in res]),
                        burn=np.int64(6000), T=np.int64(10000), seed0=np.int64(seed0))
    mk = np.array([r[0] for r in res])
    print("%-28s %d chains: mean k = %.3f +- %.3f, accept = %.3f" % (
        name, nchains, mk.mean(), mk.std() / np.sqrt(nchains), np.mean([r[2] for r in res])))


def _cp_marginal_worker(seed):
    """Thinned samples (every 400th state of steps 6000..10000) of one reference chain: sigma, k and the
    predicted step height at two query points -- the raw material of the two-sample KS gate."""
    R = refshim.load_reference()
    pm, pprop, ptheta0, _ = port.make_changepoint_problem()
    model = R.ChangepointRegression1D(pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb, pm.kmax, pm.alpha, pm.beta)
    prop = R.ChangepointRegression1DProp(model, pprop.hscale)
    np.random.seed(seed)
    s = R.Sampler(model, prop, R.ChangepointParams(ptheta0.cpx, ptheta0.cpv, ptheta0.sig))
    with refshim.quiet(), np.errstate(all="ignore"):
        s.run(10000, 6000, 400)
    xq = np.array([pm.xmin + (pm.xmax - pm.xmin) * (q + 0.5) / 6.0 for q in (1, 4)])
    sig = np.array([float(np.squeeze(t.sig)) for t in s._chain_thetas])
    ks = np.array([len(t.cpx) for t in s._chain_thetas])
    yq = np.array([np.asarray(t.cpv)[np.searchsorted(np.asarray(t.cpx), xq)] for t in s._chain_thetas])
    return sig, ks, yq


def _changepoint_marginal_samples(name, nchains=96, seed0=7000):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(_cp_marginal_worker, [seed0 + i for i in range(nchains)])
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"),
                        sig=np.array([r[0] for r in res]), k=np.array([r[1] for r in res]),
                        yq=np.array([r[2] for r in res]), query=np.array([1, 4]),
                        burn=np.int64(6000), T=np.int64(10000), thin=np.int64(400), seed0=np.int64(seed0))
    print("%-28s %d chains x %d thinned samples" % (name, nchains, len(res[0][0])))


def _n1_dense_fixtures(R):
    """"next" row N1 on the dense (d > 8) device path: leapfrog with Nsteps > 1 (hamiltonian.py:13-52),
    fixed and adaptive step size, on a d = 12 Gaussian with a dense covariance and on benchmarks.py:25-26."""
    rng = np.random.Generator(np.random.Philox(7))
    rng.standard_normal((5, 5)); rng.standard_normal(5)              # same stream position as main()
    A12 = rng.standard_normal((12, 12))
    C12 = A12 @ A12.T / 12 + 0.2 * np.eye(12)
    mu12 = rng.standard_normal(12)
    g12 = R.MultiGaussianDist(mu12, C12)
    _vector_fixture(R, "hmc4_gauss12d", g12, R.VanillaHMC(0.15, 4, g12.grad_log_likelihood), mu12 + 0.4, 500, 401,
                    extra=dict(eps=np.float64(0.15), nsteps=np.int64(4), mu=mu12, C=C12))
    _vector_fixture(R, "adapthmc3_gauss12d", g12, R.AdaptScaleHMC(0.1, 3, g12.grad_log_likelihood), mu12 - 0.3, 800, 402,
                    extra=dict(eps=np.float64(0.1), nsteps=np.int64(3), mu=mu12, C=C12), track_scale=True)
    # "next" row N2 on the dense path: pCN (randomwalk.py:78-100) with a dense covariance
    g12z = R.MultiGaussianDist(np.zeros(12), C12)                  # pCN shrinks towards 0: zero-mean target
    _vector_fixture(R, "pcn_gauss12d", g12z, R.pCN(0.6 * C12, 0.85), np.full(12, 0.3), 600, 404,
                    extra=dict(C0=0.6 * C12, rho=np.float64(0.85), mu=np.zeros(12), C=C12))
    # leapfrog with a mass matrix (hamiltonian.py:18-21, 76-91) on the dense path: M close to the target precision
    M12 = np.linalg.inv(C12) + 0.3 * np.diag(np.arange(1, 13) / 6.0)
    M12 = 0.5 * (M12 + M12.T)
    _vector_fixture(R, "hmcmass3_gauss12d", g12, R.VanillaHMC(0.35, 3, g12.grad_log_likelihood, M=M12), mu12 + 0.4, 500, 405,
                    extra=dict(eps=np.float64(0.35), nsteps=np.int64(3), mu=mu12, C=C12, M=M12))
    _vector_fixture(R, "adaptmalamass_gauss12d", g12, R.AdaptScaleHMC(0.3, 1, g12.grad_log_likelihood, M=M12), mu12 - 0.3, 800, 406,
                    extra=dict(eps=np.float64(0.3), nsteps=np.int64(1), mu=mu12, C=C12, M=M12), track_scale=True)
    g100 = R.benchmarks.benchmark_gauss100d_corr
    _vector_fixture(R, "hmc5_gauss100d", g100, R.VanillaHMC(0.1, 5, g100.grad_log_likelihood), np.zeros(100), 200, 403,
                    extra=dict(eps=np.float64(0.1), nsteps=np.int64(5)))


def _pt_fixture(R, name, model, proposal, theta0, T, seed, extra=None):
    """"next" row N3: the reference's PTSampler (default ladder 0.5**arange(5), Pswap = 0.1) with every draw
    recorded.  Per PT step and chain i the stream is: one selection uniform; then, unless the chain was swapped
    by its upper neighbour, either a within-chain step (normal(d), accept uniform) or one swap uniform."""
    np.random.seed(seed)
    with refshim.quiet():
        pt = R.PTSampler(model, proposal, theta0)
    nt, d = len(pt.betas), len(np.atleast_1d(theta0))
    usel, xi, u = np.zeros((T, nt)), np.zeros((T, nt, d)), np.zeros((T, nt))
    kind = np.zeros((T, nt), dtype=np.int8)            # 0 within-chain step, 1 swap initiator, 2 swapped partner
    with refshim.quiet(), refshim.RecordingRNG() as rec:
        for t in range(T):
            a = rec.mark()
            pt.sample()
            log = rec.log[a:rec.mark()]
            k, i = 0, 0
            while i < nt:
                assert log[k][0] == "uniform"
                usel[t, i] = float(log[k][1]); k += 1
                if i > 0 and kind[t, i - 1] == 1:
                    kind[t, i] = 2
                elif usel[t, i] > pt.Pswap or i == nt - 1:
                    assert log[k][0] == "normal" and log[k + 1][0] == "uniform", log[k:k + 2]
                    xi[t, i] = log[k][1]; u[t, i] = float(log[k + 1][1]); k += 2
                else:
                    assert log[k][0] == "uniform"
                    kind[t, i] = 1; u[t, i] = float(log[k][1]); k += 1
                i += 1
            assert k == len(log), (k, len(log))
    out = dict(usel=usel, xi=xi, u=u, kind=kind, betas=np.asarray(pt.betas, dtype=np.float64),
               pswap=np.float64(pt.Pswap),
               thetas=np.array([[np.atleast_1d(th) for th in s._chain_thetas] for s in pt.samplers]),   # [nt][T+1][d]
               logpost=np.array([s._chain_logpost for s in pt.samplers], dtype=np.float64), seed=np.int64(seed))
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print("%-28s T=%d nt=%d swaps proposed=%d accepted=%d" % (
        name, T, nt, int((kind == 1).sum()),
        int(sum(np.any(out["thetas"][i, 1:] != out["thetas"][i, :-1], axis=1)[kind[:, i] == 1].sum() for i in range(nt)))))


def _n3_pt_fixtures(R):
    B = R.benchmarks
    C0 = np.array([[0.5, 0.2], [0.2, 0.3]])
    _pt_fixture(R, "pt_rw_gauss2d", B.benchmark_gauss2d_corr, R.MetropolisRandomWalk(C0), np.ones(2), 1500, 601,
                extra=dict(C0=C0))
    rng = np.random.Generator(np.random.Philox(7))
    A = rng.standard_normal((5, 5))
    C5 = A @ A.T / 5 + 0.2 * np.eye(5)
    mu5 = rng.standard_normal(5)
    g5 = R.MultiGaussianDist(mu5, C5)
    _pt_fixture(R, "pt_rw_gauss5d", g5, R.MetropolisRandomWalk(0.3 * C5), np.zeros(5), 800, 602,
                extra=dict(C0=0.3 * C5, mu=mu5, C=C5))
    _n3_pt_dense_fixture(R)
    _n3_pt_logistic_fixture(R)


def _n3_pt_logistic_fixture(R):
    """The ladder on the logistic model (absent in the reference: the port's model behind the reference's Model base,
    driven by the reference's own PTSampler / TemperedModel / MetropolisRandomWalk): the prior is NOT tempered."""
    N, d = 400, 6
    X, y, theta_star, pv = port.make_logistic_problem(N, d, seed=port.SEED_BASE + 61)
    pmodel = port.LogisticRegression(X, y, pv)

    class RefLogistic(R.Model):
        def log_prior(self, theta):
            return pmodel.log_prior(theta)

        def log_likelihood(self, theta):
            return pmodel.log_likelihood(theta)

    rng = np.random.Generator(np.random.Philox(13))
    A = rng.standard_normal((d, d))
    C0 = 0.02 * (A @ A.T / d + 0.5 * np.eye(d))
    _pt_fixture(R, "pt_rw_logistic", RefLogistic(), R.MetropolisRandomWalk(C0), theta_star + 0.05, 600, 604,
                extra=dict(C0=C0, X=X, y=y, prior_var=np.float64(pv)))


def _n3_pt_dense_fixture(R):
    """The same ladder on the dense-Gaussian device path (d = 12 > 8): dense proposal covariance."""
    rng = np.random.Generator(np.random.Philox(11))
    A = rng.standard_normal((12, 12))
    C12 = A @ A.T / 12 + 0.2 * np.eye(12)
    mu12 = rng.standard_normal(12)
    g12 = R.MultiGaussianDist(mu12, C12)
    _pt_fixture(R, "pt_rw_gauss12d", g12, R.MetropolisRandomWalk(0.15 * C12), np.zeros(12), 700, 603,
                extra=dict(C0=0.15 * C12, mu=mu12, C=C12))


def _n5_adaptcov_fixtures(R):
    """"next" row N5: AdaptCovRandomWalk (randomwalk.py:40-56, adaptive.py:38-103) -- needs shim 6 (np.float)."""
    B = R.benchmarks
    def run(name, model, d, theta0, T, seed, **kw):
        C0 = 0.1 * np.eye(d)
        prop = R.AdaptCovRandomWalk(C0.copy(), **kw)
        _vector_fixture(R, name, model, prop, theta0, T, seed,
                        extra=dict(C0=C0, L_final=np.array(prop.L), t_adapt=np.float64(kw.get("t_adapt", 1)),
                                   marginalize=np.int64(kw.get("marginalize", False)),
                                   smooth_adapt=np.int64(kw.get("smooth_adapt", False))))
        g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        out = {k: g[k] for k in g.files}
        out["L_final"] = np.array(prop.L)                       # after the run
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    run("adaptcov_gauss2d", B.benchmark_gauss2d_corr, 2, np.ones(2), 1200, 701)
    run("adaptcov_smooth_gauss2d", B.benchmark_gauss2d_corr, 2, np.ones(2), 800, 702, t_adapt=50, smooth_adapt=True)
    rng = np.random.Generator(np.random.Philox(7))
    A = rng.standard_normal((5, 5))
    C5 = A @ A.T / 5 + 0.2 * np.eye(5)
    mu5 = rng.standard_normal(5)
    g5 = R.MultiGaussianDist(mu5, C5)
    for nm, kw in (("adaptcov_gauss5d", {}), ("adaptcov_marg_gauss5d", dict(marginalize=True)),
                   ("adaptcov_smooth_gauss5d", dict(t_adapt=30, smooth_adapt=True)),
                   ("adaptcov_smooth_marg_gauss5d", dict(t_adapt=30, smooth_adapt=True, marginalize=True))):
        run(nm, g5, 5, np.zeros(5), 1000, 703, **kw)
        g = np.load(os.path.join(GOLDEN_DIR, nm + ".npz"))
        out = {k: g[k] for k in g.files}
        out.update(mu=mu5, C=C5)
        np.savez_compressed(os.path.join(GOLDEN_DIR, nm + ".npz"), **out)


def _example_proposal_fixtures(R):
    """The remaining proposals examples/test_randomwalk.py:23-37 lists (BASELINE config 0 runs the last one)."""
    g2 = R.benchmarks.benchmark_gauss2d_corr
    # :36-38 -- the proposal the script actually runs
    p = R.AdaptScaleCovHMC(0.1, 5, g2.grad_log_likelihood, np.eye(2), t_adapt=100, smooth_adapt=True)
    p.scale = 1.0
    _vector_fixture(R, "adaptscalecovhmc5_gauss2d", g2, p, np.ones(2), 1500, 801,
                    extra=dict(eps=np.float64(0.1), nsteps=np.int64(5), M0=np.eye(2)), track_scale=True)
    M2 = np.array([[2.0, 0.3], [0.3, 1.0]])
    p = R.AdaptScaleCovHMC(0.15, 3, g2.grad_log_likelihood, M2, t_adapt=100, smooth_adapt=True)
    _vector_fixture(R, "adaptscalecovhmc3_mass_gauss2d", g2, p, np.ones(2), 1000, 802,
                    extra=dict(eps=np.float64(0.15), nsteps=np.int64(3), M0=M2), track_scale=True)
    # :25 AdaptScaleCovRandomWalk (smooth: well conditioned, see the AdaptCov fixtures)
    C0 = 0.05 * np.eye(2)
    p = R.AdaptScaleCovRandomWalk(C0.copy(), t_adapt=40, smooth_adapt=True)
    _vector_fixture(R, "adaptscalecov_rw_gauss2d", g2, p, np.ones(2), 1200, 803,
                    extra=dict(C0=C0, t_adapt=np.float64(40), smooth_adapt=np.int64(1), marginalize=np.int64(0)),
                    track_scale=True)
    # :34-35 AdaptCovHMC (smooth adaptation as in the script; the mass matrix adapts inside the leapfrog)
    p = R.AdaptCovHMC(0.1, 5, g2.grad_log_likelihood, np.eye(2), t_adapt=100, smooth_adapt=True)
    _vector_fixture(R, "adaptcovhmc5_gauss2d", g2, p, np.ones(2), 1200, 805,
                    extra=dict(eps=np.float64(0.1), nsteps=np.int64(5), M0=np.eye(2), t_adapt=np.float64(100),
                               smooth_adapt=np.int64(1), marginalize=np.int64(0)))
    # :29 AdaptScalepCN
    p = R.AdaptScalepCN(np.eye(2), 0.5)
    _vector_fixture(R, "adaptscalepcn_gauss2d", g2, p, np.ones(2), 1000, 804,
                    extra=dict(C0=np.eye(2), rho=np.float64(0.5)), track_scale=True)


def _portmodel_through_reference(R, name, kind, seed):
    """Logistic / mMALA are not in the reference: run the PORT's model (and, for
    mMALA, proposal) through the reference's own Sampler.sample and VanillaHMC."""
    nsteps = 1
    if kind == "mala":
        N, d, T, eps = 500, 8, 400, 0.35
    elif kind == "hmc3":                        # "next" row N1 on the logistic family: the reference's leapfrog, 3 steps
        N, d, T, eps, nsteps = 500, 8, 300, 0.2, 3
    elif kind == "adapthmc4":                   # AdaptScaleHMC, 4 steps, d = 20 (the dense-in-d device path)
        N, d, T, eps, nsteps = 600, 20, 300, 0.1, 4
    elif kind == "hmcmass3":                    # VanillaHMC with a (dense, fixed) mass matrix, hamiltonian.py:70-89
        N, d, T, eps, nsteps = 500, 8, 300, 0.25, 3
    elif kind == "adaptmalamass":               # AdaptScaleHMC, one step, mass matrix
        N, d, T, eps, nsteps = 600, 20, 300, 0.3, 1
    else:
        N, d, T, eps = 400, 6, 300, 0.9
    X, y, theta_star, pv = port.make_logistic_problem(N, d, seed=port.SEED_BASE + 40 + d)
    pmodel = port.LogisticRegression(X, y, pv)

    class RefLogistic(R.Model):                 # the reference's Model base (model.py)
        def log_prior(self, theta):
            return pmodel.log_prior(theta)

        def log_likelihood(self, theta):
            return pmodel.log_likelihood(theta)

    model = RefLogistic()
    track = False
    extra_m = {}
    if kind in ("hmcmass3", "adaptmalamass"):
        # a mass matrix of the posterior's scale: the Fisher information at theta_star plus the prior precision, mixed
        # with a random SPD perturbation so that it is dense and not the exact metric
        rng = np.random.RandomState(seed)
        pstar = 1.0 / (1.0 + np.exp(-X.dot(theta_star)))
        F = (X * (pstar * (1 - pstar))[:, None]).T.dot(X) + np.eye(d) / pv
        A = rng.standard_normal((d, d))
        M = F / np.mean(np.diag(F)) + 0.05 * A.dot(A.T) / d
        extra_m = dict(M=M)
        if kind == "hmcmass3":
            prop = R.VanillaHMC(eps, nsteps, pmodel.grad_log_posterior, M=M)
        else:
            prop = R.AdaptScaleHMC(eps, nsteps, pmodel.grad_log_posterior, M=M)
            track = True
    elif kind in ("mala", "hmc3"):
        prop = R.VanillaHMC(eps, nsteps, pmodel.grad_log_posterior)     # reference proposal
    elif kind == "adapthmc4":
        prop = R.AdaptScaleHMC(eps, nsteps, pmodel.grad_log_posterior)
        track = True
    else:
        class RefMMALA(R.Proposal):
            inner = port.SimplifiedMMALA(eps, pmodel)

            def propose(self, theta):
                return self.inner.propose(theta)
        prop = RefMMALA()
    theta0 = theta_star + 0.05 * np.ones(d)
    _vector_fixture(R, name, model, prop, theta0, T, seed,
                    extra=dict(X=X, y=y, prior_var=np.float64(pv), eps=np.float64(eps), nsteps=np.int64(nsteps), **extra_m),
                    track_scale=track)


def main():
    if not refshim.reference_available():
        sys.exit("reference tree not available; fixtures can only be generated in "
                 "the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if "--only-cp-marginals" in sys.argv:          # (slow fixtures can be regenerated on their own)
        _changepoint_marginal_samples("changepoint_marginals")
        return
    R = refshim.load_reference()
    if "--only-n1-dense" in sys.argv:
        _n1_dense_fixtures(R)
        return
    if "--only-n5-adaptcov" in sys.argv:
        _n5_adaptcov_fixtures(R)
        return
    if "--only-example-proposals" in sys.argv:
        _example_proposal_fixtures(R)
        return
    if "--only-n1-logistic" in sys.argv:
        _portmodel_through_reference(R, "hmc3_logistic", "hmc3", 503)
        _portmodel_through_reference(R, "adapthmc4_logistic", "adapthmc4", 504)
        return
    if "--only-mass-logistic" in sys.argv:
        _portmodel_through_reference(R, "hmcmass3_logistic", "hmcmass3", 505)
        _portmodel_through_reference(R, "adaptmalamass_logistic", "adaptmalamass", 506)
        return
    if "--only-n3-pt" in sys.argv:
        _n3_pt_fixtures(R)
        return
    if "--only-n3-pt-dense" in sys.argv:
        _n3_pt_dense_fixture(R)
        return
    if "--only-n3-pt-logistic" in sys.argv:
        _n3_pt_logistic_fixture(R)
        return
    B = R.benchmarks

    # --- config 1 family: RW on the Gaussian benchmarks (sampler.py:72-90,
    #     randomwalk.py:21-26, gaussian.py:49-52)
    C0 = np.array([[0.5, 0.2], [0.2, 0.3]])
    _vector_fixture(R, "rw_gauss2d", B.benchmark_gauss2d_corr,
                    R.MetropolisRandomWalk(C0), np.ones(2), 2000, 101, extra=dict(C0=C0))
    _vector_fixture(R, "rw_gauss1d", B.benchmark_gauss1d,
                    R.MetropolisRandomWalk(2.0), np.ones(1), 500, 102, extra=dict(C0=np.array([[2.0]])))
    _vector_fixture(R, "adaptrw_gauss2d", B.benchmark_gauss2d_corr,
                    R.AdaptScaleRandomWalk(1e-4 * np.eye(2)), np.ones(2), 3000, 103,
                    extra=dict(C0=1e-4 * np.eye(2)), track_scale=True)
    rng = np.random.Generator(np.random.Philox(7))
    A = rng.standard_normal((5, 5))
    C5 = A @ A.T / 5 + 0.2 * np.eye(5)
    mu5 = rng.standard_normal(5)
    g5 = R.MultiGaussianDist(mu5, C5)
    _vector_fixture(R, "rw_gauss5d", g5, R.MetropolisRandomWalk(0.3 * C5), np.zeros(5), 1000, 104,
                    extra=dict(C0=0.3 * C5, mu=mu5, C=C5))

    # --- MALA == VanillaHMC(eps, 1, grad)  (hamiltonian.py:55-91)
    g2 = B.benchmark_gauss2d_corr
    _vector_fixture(R, "mala_gauss2d", g2, R.VanillaHMC(0.25, 1, g2.grad_log_likelihood),
                    np.ones(2), 1000, 201, extra=dict(eps=np.float64(0.25)))
    _vector_fixture(R, "mala_gauss5d", g5, R.VanillaHMC(0.3, 1, g5.grad_log_likelihood),
                    np.zeros(5), 1000, 202, extra=dict(eps=np.float64(0.3), mu=mu5, C=C5))
    g100 = B.benchmark_gauss100d_corr
    _vector_fixture(R, "mala_gauss100d", g100, R.VanillaHMC(0.12, 1, g100.grad_log_likelihood),
                    np.zeros(100), 300, 203, extra=dict(eps=np.float64(0.12)))
    g1000 = R.MultiGaussianDist(np.zeros(1000), 0.1 * np.eye(1000) + 0.9 * np.ones((1000, 1000)))
    th0 = np.random.Generator(np.random.Philox(11)).standard_normal(1000) * 0.3
    _vector_fixture(R, "mala_gauss1000d", g1000, R.VanillaHMC(0.08, 1, g1000.grad_log_likelihood),
                    th0, 8, 204, extra=dict(eps=np.float64(0.08)))
    # RW at d=100 (benchmarks.py:25-26)
    _vector_fixture(R, "rw_gauss100d", g100, R.MetropolisRandomWalk(0.002 * np.eye(100)),
                    np.zeros(100), 300, 205, extra=dict(C0=0.002 * np.eye(100)))

    # d = 12: exercises the dense (d > 8) device path with a dense proposal covariance,
    # a non-zero mean and per-chain step-size adaptation
    A12 = rng.standard_normal((12, 12))
    C12 = A12 @ A12.T / 12 + 0.2 * np.eye(12)
    mu12 = rng.standard_normal(12)
    g12 = R.MultiGaussianDist(mu12, C12)
    _vector_fixture(R, "rw_dense_gauss12d", g12, R.MetropolisRandomWalk(0.15 * C12), mu12 + 0.5, 600, 206,
                    extra=dict(C0=0.15 * C12, mu=mu12, C=C12))
    _vector_fixture(R, "adaptrw_gauss12d", g12, R.AdaptScaleRandomWalk(0.01 * np.eye(12)), mu12 + 0.5, 800, 207,
                    extra=dict(C0=0.01 * np.eye(12), mu=mu12, C=C12), track_scale=True)
    _vector_fixture(R, "adaptmala_gauss12d", g12, R.AdaptScaleHMC(0.2, 1, g12.grad_log_likelihood),
                    mu12 - 0.3, 800, 208, extra=dict(eps=np.float64(0.2), nsteps=np.int64(1), mu=mu12, C=C12),
                    track_scale=True)

    # --- HMC proper ("next" row N1): Nsteps>1, mass matrix, adaptive scale
    _vector_fixture(R, "hmc5_gauss2d", g2, R.VanillaHMC(0.1, 5, g2.grad_log_likelihood),
                    np.ones(2), 1000, 301, extra=dict(eps=np.float64(0.1), nsteps=np.int64(5)))
    M2 = np.array([[2.0, 0.3], [0.3, 1.0]])
    _vector_fixture(R, "hmc3_mass_gauss2d", g2, R.VanillaHMC(0.2, 3, g2.grad_log_likelihood, M=M2),
                    np.ones(2), 1000, 302, extra=dict(eps=np.float64(0.2), nsteps=np.int64(3), M=M2))
    _vector_fixture(R, "mala_mass_gauss5d", g5, R.VanillaHMC(0.3, 1, g5.grad_log_likelihood,
                                                            M=np.linalg.inv(C5)),
                    np.zeros(5), 500, 303,
                    extra=dict(eps=np.float64(0.3), M=np.linalg.inv(C5), mu=mu5, C=C5))
    _vector_fixture(R, "adapthmc5_gauss2d", g2, R.AdaptScaleHMC(0.1, 5, g2.grad_log_likelihood),
                    np.ones(2), 1500, 304, extra=dict(eps=np.float64(0.1), nsteps=np.int64(5)),
                    track_scale=True)
    _n1_dense_fixtures(R)
    _n3_pt_fixtures(R)
    _n5_adaptcov_fixtures(R)
    _example_proposal_fixtures(R)
    # pCN ("next" row N2)
    _vector_fixture(R, "pcn_gauss2d", g2, R.pCN(np.eye(2), 0.5), np.ones(2), 800, 305,
                    extra=dict(C0=np.eye(2), rho=np.float64(0.5)))

    # --- config 2: changepoint model + 4-way proposal
    _changepoint_fixture(R, "changepoint", nchains=4, T=3000, seed0=400)
    _changepoint_posterior_summary("changepoint_posterior")
    _changepoint_marginal_samples("changepoint_marginals")

    # --- models/proposals the reference lacks, driven through the reference Sampler
    _portmodel_through_reference(R, "mala_logistic", "mala", 501)
    _portmodel_through_reference(R, "mmala_logistic", "mmala", 502)
    _portmodel_through_reference(R, "hmc3_logistic", "hmc3", 503)
    _portmodel_through_reference(R, "adapthmc4_logistic", "adapthmc4", 504)
    _portmodel_through_reference(R, "hmcmass3_logistic", "hmcmass3", 505)
    _portmodel_through_reference(R, "adaptmalamass_logistic", "adaptmalamass", 506)


if __name__ == "__main__":
    main()
