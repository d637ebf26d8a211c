"""
GPU parity for the logistic-regression family (riemann_b200/csrc/logistic.cu): MALA
(config 4 shape) and simplified manifold MALA (config 5 shape).

The reference has neither the model nor the metric proposal ("parity unpinned"):
the fixtures were produced by driving the numpy restatement of the model through the
REFERENCE's own Sampler.sample and VanillaHMC (oracle/gen_golden.py), and the device
replays that stream.

Tolerance: fp64 kernels vs fp64 numpy.  Log-posteriors 1e-9 relative-or-absolute, gradients
and metric 1e-9, mMALA log q ratio through the decisions (identical) and the chains (1e-8:
the Cholesky / triangular solves amplify round-off by the metric's condition number).
"""
import numpy as np
import pytest

from gpu_helpers import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _models(X, y, pv):
    from oracle import riemann_port as port
    from riemann_b200.models.logistic import LogisticRegression
    return LogisticRegression(X, y, pv), port.LogisticRegression(X, y, pv)


@pytest.mark.parametrize("N,d", [(500, 8), (1000, 100), (333, 7), (400, 6), (5000, 64), (130, 1)])
def test_pointwise_logpost_grad_metric(N, d):
    from oracle import riemann_port as port
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=100 + d)
    dm, om = _models(X, y, pv)
    rng = np.random.default_rng(d)
    Th = ts[None, :] * rng.uniform(0, 1.5, (13, 1)) + 0.3 * rng.standard_normal((13, d))
    assert relerr(dm.log_posterior_batch(Th).cpu().numpy(), [om.log_posterior(t) for t in Th]) < TOL
    assert relerr(dm.log_likelihood_batch(Th).cpu().numpy(), [om.log_likelihood(t) for t in Th]) < TOL
    assert relerr(dm.log_prior_batch(Th).cpu().numpy(), [om.log_prior(t) for t in Th]) < TOL
    assert relerr(dm.grad_log_posterior_batch(Th).cpu().numpy(), [om.grad_log_posterior(t) for t in Th]) < TOL
    if d <= 64:
        assert relerr(dm.metric_batch(Th).cpu().numpy(), [om.metric(t) for t in Th]) < TOL
    assert abs(dm.log_posterior(Th[0]) - om.log_posterior(Th[0])) < TOL * abs(om.log_posterior(Th[0]))


def test_extreme_logits_are_stable():
    from oracle import riemann_port as port
    X, y, ts, pv = port.make_logistic_problem(300, 5, seed=3)
    dm, om = _models(X, y, pv)
    Th = np.stack([ts * 400.0, -ts * 400.0, np.zeros(5)])            # |z| up to several hundred
    assert relerr(dm.log_posterior_batch(Th).cpu().numpy(), [om.log_posterior(t) for t in Th]) < TOL
    assert relerr(dm.grad_log_posterior_batch(Th).cpu().numpy(), [om.grad_log_posterior(t) for t in Th]) < 1e-8


@pytest.mark.parametrize("name,tol", [("mala_logistic", 1e-9), ("mmala_logistic", 1e-8),
                                      # "next" row N1 on this family: the reference's leapfrog with Nsteps > 1
                                      ("hmc3_logistic", 1e-9), ("adapthmc4_logistic", 1e-9),
                                      # ... and with a fixed dense mass matrix (hamiltonian.py:70-89)
                                      ("hmcmass3_logistic", 1e-9), ("adaptmalamass_logistic", 1e-9)])
def test_injected_chain_matches_fixture(golden, name, tol):
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA, VanillaHMC, AdaptScaleHMC
    g = golden(name)
    dm, _ = _models(g["X"], g["y"], float(g["prior_var"]))
    if name == "mala_logistic":
        p = MALA(float(g["eps"]), dm.grad_log_posterior)
    elif name == "hmc3_logistic":
        p = VanillaHMC(float(g["eps"]), int(g["nsteps"]), dm.grad_log_posterior)
    elif name == "adapthmc4_logistic":
        p = AdaptScaleHMC(float(g["eps"]), int(g["nsteps"]), dm.grad_log_posterior)
    elif name == "hmcmass3_logistic":
        p = VanillaHMC(float(g["eps"]), int(g["nsteps"]), dm.grad_log_posterior, M=g["M"])
    elif name == "adaptmalamass_logistic":
        p = AdaptScaleHMC(float(g["eps"]), int(g["nsteps"]), dm.grad_log_posterior, M=g["M"])
    else:
        p = SimplifiedMMALA(float(g["eps"]), dm)
    s = Sampler(dm, p, g["thetas"][0])
    ex = s.run_injected(xi=g["xi"], u=g["u"])
    assert relerr(np.array(s._chain_thetas), g["thetas"]) < tol
    assert relerr(s._chain_logpost, g["logpost"]) < tol
    assert relerr(ex["prop_logpost"][:, 0], g["prop_logpost"]) < tol
    assert np.array_equal(ex["accepted"][:, 0], np.any(g["thetas"][1:] != g["thetas"][:-1], axis=1))


@pytest.mark.parametrize("kind", ["mala", "mmala"])
def test_ragged_chains_vs_oracle(kind):
    """K = 70 chains (not a multiple of the 64-chain CTA tile), several row splits."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA
    N, d, K, T = 1500, 10, 70, 25
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=77)
    dm, om = _models(X, y, pv)
    rng = np.random.default_rng(1)
    th0 = ts[None, :] + 0.1 * rng.standard_normal((K, d))
    xi = rng.standard_normal((T, K, d))
    u = rng.uniform(size=(T, K))
    eps = 0.15 if kind == "mala" else 0.7
    p = MALA(eps, dm.grad_log_posterior) if kind == "mala" else SimplifiedMMALA(eps, dm)
    s = Sampler(dm, p, th0)
    s.run_injected(xi=xi, u=u)
    for c in (0, 63, 64, 69):
        op = port.MALA(eps, om.grad_log_posterior) if kind == "mala" else port.SimplifiedMMALA(eps, om)
        o = port.Sampler(om, op, th0[c], draws=port.VectorTapeDraws(xi[:, c], u[:, c]))
        o.run(T)
        assert relerr(s._chain_thetas[:, c], np.array(o._chain_thetas)) < 1e-8
        assert relerr(s._chain_logpost[:, c], np.array(o._chain_logpost)) < 1e-8
    acc = np.mean(np.any(s._chain_thetas[1:] != s._chain_thetas[:-1], axis=2))
    assert 0.2 < acc < 0.99


@pytest.mark.parametrize("kind,prec", [("rw", "f64"), ("adaptrw", "f64"), ("pcn", "f64"), ("rw", "tf32x3"), ("pcn", "tf32x3")])
def test_random_walk_and_pcn_on_the_logistic_model(kind, prec):
    """MetropolisRandomWalk / AdaptScaleRandomWalk (riemann/proposals/randomwalk.py:12-37) and pCN (:78-100) with a
    DENSE proposal covariance on the logistic model: the oracle's classes driven by the reference's accept rule on the
    same injected stream.  f64: chains to 1e-8; tf32x3: same decisions except inside the stated difference budget."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler, budgets
    from riemann_b200.proposals.randomwalk import AdaptScaleRandomWalk, MetropolisRandomWalk, pCN
    N, d, K, T = 2500, 12, 37, 30
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=91)
    dm, om = _models(X, y, pv)
    rng = np.random.default_rng(6)
    A = rng.standard_normal((d, d))
    C = 2e-3 * (A @ A.T / d + 0.5 * np.eye(d))
    th0 = ts[None, :] + 0.05 * rng.standard_normal((K, d))
    xi, u = rng.standard_normal((T, K, d)), rng.uniform(size=(T, K))
    if kind == "pcn":
        C = 200.0 * C                    # pCN is reversible for N(0, C): a wide C keeps rho theta near the mode
        mk_d, mk_o = (lambda: pCN(C, 0.999)), (lambda: port.pCN(C, 0.999))
    elif kind == "rw":
        mk_d, mk_o = (lambda: MetropolisRandomWalk(C)), (lambda: port.MetropolisRandomWalk(C))
    else:
        mk_d, mk_o = (lambda: AdaptScaleRandomWalk(C)), (lambda: port.AdaptScaleRandomWalk(C))
    s = Sampler(dm, mk_d(), th0, precision=prec)
    ex = s.run_injected(xi=xi, u=u)
    n_diff = 0
    for c in (0, 5, 31, 36):
        o = port.Sampler(om, mk_o(), th0[c], draws=port.VectorTapeDraws(xi[:, c], u[:, c]))
        o.run(T)
        oth, olp = np.array(o._chain_thetas), np.array(o._chain_logpost)
        if prec == "f64":
            assert relerr(s._chain_thetas[:, c], oth) < 1e-8
            assert relerr(s._chain_logpost[:, c], olp) < 1e-8
        else:
            same = np.all(np.abs(s._chain_thetas[:, c] - oth) < 1e-9, axis=1)
            n_diff += int(not same.all())
            k = int(np.argmin(same)) if not same.all() else T + 1       # first step the two chains part (a close call)
            assert np.max(np.abs(s._chain_logpost[:k, c] - olp[:k])) < budgets.logistic_tf32x3_offset(N)
    assert n_diff <= 1
    acc = ex["accepted"].mean()
    assert 0.05 < acc < 0.98, acc
    if kind == "adaptrw":
        o = port.Sampler(om, mk_o(), th0[0], draws=port.VectorTapeDraws(xi[:, 0], u[:, 0]))
        o.run(T)
        assert abs(np.atleast_1d(s.proposal.scale)[0] / o.proposal.scale - 1) < 1e-9


@pytest.mark.parametrize("kind", ["mala", "mmala"])
def test_philox_run_agrees_with_oracle_posterior(kind):
    """Distributional gate vs a long CPU chain of the oracle (same model, own numpy stream)."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA
    N, d = 400, 5
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=21)
    dm, om = _models(X, y, pv)
    eps = 0.3 if kind == "mala" else 0.9
    np.random.seed(5)
    op = port.MALA(eps, om.grad_log_posterior) if kind == "mala" else port.SimplifiedMMALA(eps, om)
    o = port.Sampler(om, op, ts.copy())
    o.run(6000, 1000)
    och = np.array(o._chain_thetas)
    p = MALA(eps, dm.grad_log_posterior) if kind == "mala" else SimplifiedMMALA(eps, dm)
    s = Sampler(dm, p, ts.copy(), K=2048, seed=17)
    s.run(300, trace=False)
    s.reset_diagnostics()
    s.run(300, trace=False)
    dg = s.diagnostics(allreduce=False)
    th = np.asarray(s._chain_thetas[-1])
    sd = och.std(0)
    assert np.all(np.abs(th.mean(0) - och.mean(0)) < 0.25 * sd)
    assert np.all(np.abs(th.std(0) / sd - 1.0) < 0.2)
    oacc = np.mean(np.any(och[1:] != och[:-1], axis=1))
    assert abs(dg["accept_rate"] - oacc) < 0.06
    lp = np.asarray(s._chain_logpost[-1])
    assert relerr(lp, dm.log_posterior_batch(th).cpu().numpy()) < 1e-10


def test_config5_shape_self_consistency():
    """N = 1e5, d = 64 (config 5 data shape), 128 chains of mMALA: carried log-posterior equals
    a fresh evaluation; the metric is symmetric positive definite."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import SimplifiedMMALA
    X, y, ts, pv = port.make_logistic_problem(100000, 64)
    dm, om = _models(X, y, pv)
    rng = np.random.default_rng(0)
    th0 = ts[None, :] + 0.01 * rng.standard_normal((128, 64))
    s = Sampler(dm, SimplifiedMMALA(0.5, dm), th0, seed=1)
    s.run(4, trace=False)
    th, lp = s.state_tensors()
    assert relerr(lp.cpu().numpy(), dm.log_posterior_batch(th).cpu().numpy()) < 1e-10
    G = dm.metric_batch(th[:3]).cpu().numpy()
    assert relerr(G[0], om.metric(th[0].cpu().numpy())) < TOL
    assert np.allclose(G, np.transpose(G, (0, 2, 1)), rtol=1e-12, atol=1e-12)
    assert np.all(np.linalg.eigvalsh(G[1]) > 0)
    dg = s.diagnostics(allreduce=False)
    assert dg["accept_rate"] > 0.3


# ---------------------------------------------------------------------------------------
# precision="tf32-metric": the Fisher metric of the mMALA proposal as one tcgen05 TF32 GEMM
# ---------------------------------------------------------------------------------------
def test_tf32_metric_is_a_deterministic_function_of_theta():
    """Detailed balance needs G~(theta) to be ONE function of theta: chains holding the same state and
    fed the same noise must produce bit-identical proposals wherever they sit in the GEMM's tiling
    (K = 300 spans three 128-row tiles), and repeated launches must agree bit for bit."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import SimplifiedMMALA
    N, d = 3000, 20
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=4)
    dm, _ = _models(X, y, pv)
    K, T = 300, 4
    rng = np.random.default_rng(1)
    xi = np.repeat(rng.standard_normal((T, 1, d)), K, axis=1)
    u = np.repeat(rng.uniform(size=(T, 1)), K, axis=1)
    runs = []
    for _ in range(2):
        s = Sampler(dm, SimplifiedMMALA(0.7, dm), np.repeat(ts[None], K, 0), precision="tf32-metric")
        ex = s.run_injected(xi=xi, u=u)
        th = np.asarray(s._chain_thetas)                       # [T+1][K][d]
        assert np.all(th == th[:, :1]) and np.all(ex["prop_logpost"] == ex["prop_logpost"][:, :1])
        runs.append((th.copy(), np.asarray(s._chain_logpost).copy(), ex["accepted"].copy()))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    assert runs[0][2].any()                                   # the chain moves


@pytest.mark.parametrize("prec", ["tf32-metric", "tf32x3"])
def test_metric_gemm_split_k_matches_the_unsplit_product(prec, monkeypatch):
    """A small chain shard leaves most SMs without a tile of the Fisher-metric GEMM, so its contraction over the data
    rows is split into partial products that the Cholesky kernel sums (RMN_MMALA_KSPLIT forces a count; the default
    picks about one wave of tiles).  Same states, same noise: proposals with 1, 3 and the default number of partials
    differ only by the fp32 summation order (<< the metric's own rounding), and every setting is deterministic."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import SimplifiedMMALA
    N, d = 6000, 24
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=12)
    dm, _ = _models(X, y, pv)
    K = 200
    rng = np.random.default_rng(6)
    th0 = ts[None] + 0.05 * rng.standard_normal((K, d))
    xi, u = rng.standard_normal((2, K, d)), np.ones((2, K))              # u = 1: nothing is accepted
    out = {}
    for split in ("1", "3", None, "3"):
        if split is None:
            monkeypatch.delenv("RMN_MMALA_KSPLIT", raising=False)
        else:
            monkeypatch.setenv("RMN_MMALA_KSPLIT", split)
        s = Sampler(dm, SimplifiedMMALA(0.8, dm), th0, precision=prec)
        ex = s.run_injected(xi=xi, u=u)
        if split in out:
            assert np.array_equal(out[split][0], ex["prop_theta"]) and np.array_equal(out[split][1], ex["logqratio"])
        out[split] = (ex["prop_theta"].copy(), ex["logqratio"].copy())
    step = np.linalg.norm(out["1"][0][0] - th0, axis=1)
    for split in ("3", None):
        assert np.all(np.linalg.norm(out[split][0][0] - out["1"][0][0], axis=1) < 1e-4 * step)
        assert np.max(np.abs(out[split][1] - out["1"][1])) < 1e-3
    assert np.any(out["3"][0] != out["1"][0])                            # the partials really were summed differently


def test_tf32_metric_proposals_track_the_fp64_metric():
    """Same states, same noise: the proposal built with the TF32 metric differs from the fp64 one by the
    metric's rounding only (<~ 1e-3 relative in G, so ~1e-3 of the step), and its log-posterior -- still
    evaluated in fp64 -- equals a fresh fp64 evaluation at the device's own proposed point."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import SimplifiedMMALA
    N, d = 5000, 16
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=9)
    dm, om = _models(X, y, pv)
    K = 70
    rng = np.random.default_rng(2)
    th0 = ts[None] + 0.05 * rng.standard_normal((K, d))
    xi, u = rng.standard_normal((1, K, d)), np.ones((1, K))          # u = 1: nothing is accepted
    out = {}
    for prec in ("f64", "tf32-metric"):
        s = Sampler(dm, SimplifiedMMALA(0.8, dm), th0, precision=prec)
        out[prec] = s.run_injected(xi=xi, u=u)
    pa, pb = out["f64"]["prop_theta"][0], out["tf32-metric"]["prop_theta"][0]
    step = np.linalg.norm(pa - th0, axis=1)
    assert np.all(np.linalg.norm(pa - pb, axis=1) < 5e-3 * step)
    assert np.any(pa != pb)
    want = np.array([om.log_posterior(t) for t in pb])
    assert relerr(out["tf32-metric"]["prop_logpost"][0], want) < 1e-9
    assert np.max(np.abs(out["tf32-metric"]["logqratio"][0] - out["f64"]["logqratio"][0])) < 0.05


def test_tf32_metric_chain_targets_the_same_posterior():
    """Distributional gate: KS of every marginal against a long thinned oracle chain (fp64 metric)."""
    from scipy import stats
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import SimplifiedMMALA
    N, d = 400, 5
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=21)
    dm, om = _models(X, y, pv)
    np.random.seed(5)
    o = port.Sampler(om, port.SimplifiedMMALA(0.9, om), ts.copy())
    o.run(13000, 1000, 20)
    och = np.array(o._chain_thetas)
    acc = {}
    for prec in ("tf32-metric", "f64"):
        s = Sampler(dm, SimplifiedMMALA(0.9, dm), ts.copy(), K=2048, seed=23, precision=prec)
        s.run(300, trace=False)
        s.reset_diagnostics()
        s.run(300, trace=False)
        acc[prec] = s.diagnostics(allreduce=False)["accept_rate"]
        th = np.asarray(s._chain_thetas[-1])
        for j in range(d):
            assert stats.ks_2samp(th[:, j], och[:, j]).pvalue > 1e-3
        if prec == "tf32-metric":
            th_t, lp_t = th, np.asarray(s._chain_logpost[-1])
    assert abs(acc["tf32-metric"] - acc["f64"]) < 0.01           # the rounded metric costs no acceptance
    th, lp = th_t, lp_t
    assert relerr(lp, dm.log_posterior_batch(th).cpu().numpy()) < 1e-10


def test_tf32_metric_rejects_other_models():
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    from riemann_b200.sampling_errors import ParameterError
    with pytest.raises((ParameterError, RuntimeError)):
        Sampler(benchmarks.benchmark_gauss2d_corr, MetropolisRandomWalk(np.eye(2)), np.ones(2), precision="tf32-metric")


def test_fast_sigmoid_softplus_accuracy():
    """logistic_math.cuh (table-driven fp64 exp / log / reciprocal) against 80-bit long double."""
    import torch
    from riemann_b200 import _lib
    rng = np.random.default_rng(0)
    z = np.concatenate([rng.standard_normal(200000) * 3, rng.uniform(-70, 70, 100000), np.linspace(-1e-3, 1e-3, 2001),
                        [0.0, -0.0, 1e-300, -1e-300, 36.0, -36.0, 64.0, -64.0, 700.0, -700.0, 1e6, -1e6, np.inf, -np.inf]])
    dz = torch.as_tensor(z, device="cuda")
    out = [torch.empty_like(dz) for _ in range(3)]
    _lib.check(_lib.load().rmn_logistic_math(len(z), _lib.ptr(dz), *[_lib.ptr(o) for o in out], _lib.stream_ptr()))
    p, sp, pq = (o.cpu().numpy() for o in out)
    zl = z.astype(np.longdouble)
    with np.errstate(all="ignore"):
        el = np.exp(-np.abs(zl))
        p_ref = np.where(zl >= 0, 1 / (1 + el), el / (1 + el))
        sp_ref = np.maximum(zl, 0) + np.log1p(el)
        pq_ref = el / (1 + el) ** 2
    fin = np.isfinite(z)
    assert np.max(np.abs(p[fin] - p_ref[fin])) < 3e-16
    assert np.max(np.abs(sp[fin] - sp_ref[fin]) / np.maximum(1.0, np.abs(sp_ref[fin]).astype(np.float64))) < 4e-16
    assert np.max(np.abs(pq[fin] - pq_ref[fin])) < 2e-16
    small = fin & (np.abs(z) < 30)
    assert np.max(np.abs(p[small] / p_ref[small].astype(np.float64) - 1)) < 1e-15        # relative, both tails
    assert p[-2] == 1.0 and sp[-2] == np.inf and 0.0 <= p[-1] < 1e-27 and abs(sp[-1]) < 2e-16   # +-inf (|z| clamped at 64)
    dn = torch.as_tensor(np.array([np.nan]), device="cuda")
    o1 = [torch.empty_like(dn) for _ in range(3)]
    _lib.check(_lib.load().rmn_logistic_math(1, _lib.ptr(dn), *[_lib.ptr(o) for o in o1], _lib.stream_ptr()))
    assert all(np.isnan(o.cpu().numpy()[0]) for o in o1)


# ---------------------------------------------------------------------------------------
# precision="tf32x3" on the logistic model: the likelihood sweep on the tcgen05 tensor cores
# ---------------------------------------------------------------------------------------
def _decision_gate(dm, ex, lp_start, u, th_start, budget):
    """The accept test of sampler.py:83-84 re-done with FP64 log-posteriors of the device's own states: decisions must
    be identical except where log u is within `budget` of the threshold.  Returns (#steps checked, #inside the band)."""
    lp64 = dm.log_posterior_batch(th_start).cpu().numpy()
    T = ex["prop_theta"].shape[0]
    n_band = 0
    th = th_start.copy()
    for t in range(T):
        lpp64 = dm.log_posterior_batch(ex["prop_theta"][t]).cpu().numpy()
        delta = lpp64 - lp64 - ex["logqratio"][t]
        mh = np.where(delta < 0, delta, 0.0)
        with np.errstate(divide="ignore"):
            margin = np.log(u[t]) - mh
        want = margin < 0
        forced = u[t] >= 1.0                      # log u = 0 is never below min(0, delta): a forced rejection, not a close call
        assert not np.any(ex["accepted"][t][forced])
        band = (np.abs(margin) < budget) & ~forced
        assert np.array_equal(ex["accepted"][t][~band], want[~band]), "decision differs outside the budget band at step %d" % t
        n_band += int(band.sum())
        acc = ex["accepted"][t]
        th[acc] = ex["prop_theta"][t][acc]
        lp64[acc] = lpp64[acc]
    return T * th.shape[0], n_band


@pytest.mark.parametrize("kind,N,d,K", [("mala", 20000, 20, 300), ("mmala", 9000, 12, 70), ("mala", 4099, 100, 130),
                                        ("mala", 30011, 64, 257), ("mmala", 12000, 64, 129)])
def test_tf32x3_logpost_and_proposal_budget(kind, N, d, K):
    """The tensor-core sweep (fused tcgen05 kernel, logistic_fused.cu) against fp64 evaluations
    of the same states, with the budgets of riemann_b200/budgets.py: the offset of one state's log-posterior, the
    DIFFERENCE proposal - state that enters the accept test, the accept decisions themselves, the proposals against
    the fp64 sampler on the same noise, and determinism across tile positions.  Ragged shapes (N % 64, K % 128)."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler, budgets
    from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=31)
    dm, _ = _models(X, y, pv)
    rng = np.random.default_rng(3)
    th0 = ts[None] + 0.05 * rng.standard_normal((K, d))
    th0[K // 2:] = th0[0]                                             # second half: identical states
    xi = rng.standard_normal((3, K, d)); xi[:, K // 2:] = xi[:, :1]
    u = rng.uniform(size=(3, K)); u[:, K // 2:] = u[:, :1]; u[0] = 1.0
    mk = (lambda: MALA(0.05, dm.grad_log_posterior)) if kind == "mala" else (lambda: SimplifiedMMALA(0.7, dm))
    out, lp0 = {}, {}
    for prec in ("f64", "tf32x3"):
        s = Sampler(dm, mk(), th0, precision=prec)
        lp0[prec] = np.asarray(s._chain_logpost[0]).copy()
        out[prec] = s.run_injected(xi=xi, u=u)
    off, dif = budgets.logistic_tf32x3_offset(N), budgets.logistic_tf32x3_difference(N)
    e0 = lp0["tf32x3"] - lp0["f64"]
    assert np.max(np.abs(e0)) < off
    assert np.max(np.abs(e0)) > 0                                     # it IS the fp32-accurate path
    # first step (nothing accepted yet: u = 1): same states in both modes, proposals agree to the TF32 gradient
    pa, pb = out["f64"]["prop_theta"][0], out["tf32x3"]["prop_theta"][0]
    step = np.linalg.norm(pa - th0, axis=1)
    assert np.all(np.linalg.norm(pa - pb, axis=1) < 5e-3 * step)
    want = dm.log_posterior_batch(pb).cpu().numpy()
    e1 = out["tf32x3"]["prop_logpost"][0] - want
    assert np.max(np.abs(e1)) < off
    assert np.max(np.abs(e1 - e0)) < dif                              # what enters the accept test
    # accept decisions of all three steps against fp64 log-posteriors of the device's own states
    h = K // 2                                                        # (chains h.. repeat chain 0: count it once)
    sub = {k: v[:, :h] for k, v in out["tf32x3"].items()}
    n, n_band = _decision_gate(dm, sub, lp0["f64"][:h], u[:, :h], th0[:h], dif)
    assert n_band <= max(2, int(0.01 * n))
    # determinism: chains K/2.. hold the same state and noise as chain K/2
    for key in ("prop_theta", "prop_logpost", "logqratio"):
        a = out["tf32x3"][key]
        assert np.array_equal(a[:, h:], np.repeat(a[:, h:h + 1], K - h, axis=1)), key


def test_tf32x3_logistic_chain_targets_the_same_posterior():
    from scipy import stats
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import MALA
    N, d = 400, 5
    X, y, ts, pv = port.make_logistic_problem(N, d, seed=21)
    dm, om = _models(X, y, pv)
    np.random.seed(5)
    o = port.Sampler(om, port.MALA(0.3, om.grad_log_posterior), ts.copy())
    o.run(13000, 1000, 20)
    och = np.array(o._chain_thetas)
    acc = {}
    for prec in ("tf32x3", "f64"):
        s = Sampler(dm, MALA(0.3, dm.grad_log_posterior), ts.copy(), K=2048, seed=29, precision=prec)
        s.run(300, trace=False)
        s.reset_diagnostics()
        s.run(300, trace=False)
        acc[prec] = s.diagnostics(allreduce=False)["accept_rate"]
        th = np.asarray(s._chain_thetas[-1])
        for j in range(d):
            assert stats.ks_2samp(th[:, j], och[:, j]).pvalue > 1e-3
    assert abs(acc["tf32x3"] - acc["f64"]) < 0.01


def test_tf32x3_logistic_config4_budget():
    """BASELINE config 4 shape (N = 1e6, d = 100), the stated budget (riemann_b200/budgets.py): one state's offset
    < 5e-8 N, the log-posterior DIFFERENCE proposal - state within 1e-3 of fp64 (SURVEY 8d's gate), accept decisions identical to an fp64
    re-evaluation outside that band; K = 2,100 chains (17 chain blocks of the fused kernel, the last one ragged)."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler, budgets
    from riemann_b200.proposals.hamiltonian import MALA
    N, d, K = 1000000, 100, 2100
    X, y, ts, pv = port.make_logistic_problem(N, d)
    dm, _ = _models(X, y, pv)
    rng = np.random.default_rng(8)
    th0 = ts[None] + 0.01 * rng.standard_normal((K, d))
    off, dif = budgets.logistic_tf32x3_offset(N), budgets.logistic_tf32x3_difference(N)
    assert dif == 1e-3
    s = Sampler(dm, MALA(0.02, dm.grad_log_posterior), th0, precision="tf32x3", seed=1)
    lp = np.asarray(s._chain_logpost[0])
    sel = np.r_[0:160, K - 40:K]
    w0 = dm.log_posterior_batch(th0[sel]).cpu().numpy()
    assert np.max(np.abs(lp[sel] - w0)) < off
    xi, u = rng.standard_normal((2, K, d)), rng.uniform(size=(2, K))
    ex = s.run_injected(xi=xi, u=u)
    w1 = dm.log_posterior_batch(ex["prop_theta"][0][sel]).cpu().numpy()
    assert np.max(np.abs((ex["prop_logpost"][0][sel] - lp[sel]) - (w1 - w0))) < dif
    sub = {k: v[:, sel] for k, v in ex.items()}
    n, n_band = _decision_gate(dm, sub, w0, u[:, sel], th0[sel], dif)
    assert n_band <= max(2, int(0.01 * n))
    assert 0.2 < ex["accepted"].mean() <= 1.0
