"""
GPU parity for parallel tempering ("next" row N3): riemann_b200.PTSampler / rmn_sampler_set_tempering against
the recorded stream of the reference's PTSampler (riemann/samplers/ptsampler.py:41-127, fixtures
tests/golden/pt_rw_gauss{2,5}d.npz written by oracle/gen_golden.py): every temperature's chain, tempered
log-posterior and swap decision at every step; several ladders side by side; distributional checks in Philox mode.
Tolerance 1e-9 (fp64 device vs fp64 numpy).
"""
import numpy as np
import pytest

from gpu_helpers import device_gauss, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _pt(g, d, K=1, **kw):
    from riemann_b200 import PTSampler
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    return PTSampler(device_gauss(g, d), MetropolisRandomWalk(g["C0"]), g["thetas"][0, 0], K=K, **kw)


@pytest.mark.parametrize("name,d", [("pt_rw_gauss2d", 2), ("pt_rw_gauss5d", 5),
                                    ("pt_rw_gauss12d", 12)])       # d > 8: the dense-Gaussian device path
def test_injected_ladder_matches_reference(golden, name, d):
    g = golden(name)
    pt = _pt(g, d)
    assert np.array_equal(pt.betas, g["betas"]) and pt.Pswap == float(g["pswap"])
    ex = pt.run_injected(g["usel"], g["xi"], g["u"])
    T, nt = g["usel"].shape
    for i in range(nt):
        assert relerr(np.array(pt.samplers[i]._chain_thetas), g["thetas"][i]) < TOL
        assert relerr(np.array(pt.samplers[i]._chain_logpost), g["logpost"][i]) < TOL
    assert relerr(np.array(pt._chain_thetas), g["thetas"][0]) < TOL          # ptsampler.py:89-90
    # swap decisions: an accepted swap shows as both partners changing state at a swap step
    kind = g["kind"]
    moved_ref = np.any(g["thetas"][:, 1:] != g["thetas"][:, :-1], axis=2).T   # [T][nt]
    acc = ex["accepted"].astype(bool)                                         # [T][nt]
    # (a swap of two IDENTICAL states -- every rung starts at theta0 -- is accepted without anything moving)
    th = g["thetas"]                                                          # [nt][T+1][d]
    same_next = np.zeros((T, nt), dtype=bool)
    same_next[:, :-1] = np.all(th[:-1, :-1] == th[1:, :-1], axis=2).T         # rung i and i+1 equal before step t
    same_prev = np.zeros((T, nt), dtype=bool)
    same_prev[:, 1:] = same_next[:, :-1]
    assert np.array_equal(acc[kind == 1], (moved_ref | same_next)[kind == 1])
    assert np.array_equal(acc[kind == 2], (moved_ref | same_prev)[kind == 2])


def test_many_ladders_each_replay_the_stream(golden):
    """K = 13 ladders x 5 temperatures = 65 chains (ladders straddle warp boundaries in chain space, not in
    thread space): every ladder fed the same stream reproduces the reference."""
    g = golden("pt_rw_gauss2d")
    K, T = 13, 400
    pt = _pt(g, 2, K=K)
    rep = lambda a: np.concatenate([a[:T]] * K, axis=1)
    pt.run_injected(rep(g["usel"]), rep(g["xi"]), rep(g["u"]))
    th, lp = pt.samplers[3]._chain_thetas, pt.samplers[3]._chain_logpost      # [rec, K, d]
    for l in (0, 6, 12):
        assert relerr(th[:, l], g["thetas"][3, :T + 1]) < TOL
        assert relerr(lp[:, l], g["logpost"][3, :T + 1]) < TOL


def test_philox_ladder_samples_every_tempered_target(golden):
    """Philox mode, 512 ladders: temperature i must sample N(mu, C / beta_i); swap moves keep that invariant."""
    from scipy import stats
    g = golden("pt_rw_gauss2d")
    pt = _pt(g, 2, K=512, seed=5)
    pt.run(3000, trace=False)
    th = pt._thetas[-1]                                                       # [K, nt, d]
    for i, beta in enumerate(pt.betas):
        x = th[:, i]
        assert stats.kstest(x[:, 0] * np.sqrt(beta), "norm").pvalue > 1e-3
        assert stats.kstest((x[:, 0] - x[:, 1]) * np.sqrt(beta / 0.2), "norm").pvalue > 1e-3
    lp = pt._logpost[-1]
    from riemann_b200.models import benchmarks
    ll = benchmarks.benchmark_gauss2d_corr.log_posterior_batch(th.reshape(-1, 2)).cpu().numpy().reshape(512, -1)
    assert relerr(lp, ll * pt.betas[None, :]) < 1e-10                         # carried tempered log-posterior


def test_tempering_rejects_what_the_reference_cannot_mean(golden):
    from riemann_b200 import PTSampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import AdaptScaleRandomWalk, MetropolisRandomWalk
    from riemann_b200.sampling_errors import ParameterError
    m = benchmarks.benchmark_gauss2d_corr
    with pytest.raises(ParameterError):
        PTSampler(m, MetropolisRandomWalk(np.eye(2)), np.ones(2), Pswap=1.5)          # ptsampler.py:71-72
    with pytest.raises(ParameterError):
        PTSampler(m, AdaptScaleRandomWalk(np.eye(2)), np.ones(2))                    # shared adaptive proposal
    with pytest.raises(ParameterError):
        PTSampler(m, MetropolisRandomWalk(np.eye(2)), np.ones(2), betas=[1.0, 1.5])   # :26-27


def test_explicit_ladder_matches_oracle_port():
    """A 3-temperature ladder with Pswap = 0.3 (the reference's `betas=` branch raises, ptsampler.py:62; the port
    implements what it intends): device vs oracle on a random tape, every chain and decision."""
    from oracle import riemann_port as port
    from riemann_b200 import PTSampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    betas, T, d = np.array([1.0, 0.4, 0.1]), 600, 2
    rng = np.random.default_rng(12)
    usel, xi, u = rng.uniform(size=(T, 3)), rng.standard_normal((T, 3, d)), rng.uniform(size=(T, 3))
    C0 = np.array([[0.5, 0.2], [0.2, 0.3]])
    om = port.benchmark_gauss(2)
    opt = port.PTSampler(om, port.MetropolisRandomWalk(C0), np.ones(2), betas=betas, Pswap=0.3,
                         draws=port.PTTapeDraws(usel, xi, u))
    with np.errstate(all="ignore"):
        opt.run(T)
    pt = PTSampler(benchmarks.benchmark_gauss2d_corr, MetropolisRandomWalk(C0), np.ones(2), betas=betas, Pswap=0.3)
    pt.run_injected(usel, xi, u)
    for i in range(3):
        assert relerr(np.array(pt.samplers[i]._chain_thetas), np.array(opt.samplers[i]._chain_thetas)) < TOL
        assert relerr(np.array(pt.samplers[i]._chain_logpost), np.array(opt.samplers[i]._chain_logpost)) < TOL
    swaps = sum(np.any(np.array(opt.samplers[0]._chain_thetas)[1:] != np.array(opt.samplers[0]._chain_thetas)[:-1], axis=1))
    assert swaps > 50


def test_philox_ladder_with_pcn():
    """Tempering with the other supported proposal: pCN (randomwalk.py:78-100) inside the ladder."""
    from scipy import stats
    from riemann_b200 import PTSampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import pCN
    pt = PTSampler(benchmarks.benchmark_gauss2d_corr, pCN(np.array([[1.0, 0.9], [0.9, 1.0]]), 0.8), np.ones(2),
                   betas=[1.0, 0.5, 0.25, 0.125], Pswap=0.2, K=512, seed=8)
    pt.run(2000, trace=False)
    th = pt._thetas[-1]
    for i, beta in enumerate(pt.betas):
        assert stats.kstest(th[:, i, 0] * np.sqrt(beta), "norm").pvalue > 1e-3
        assert stats.kstest((th[:, i, 0] + th[:, i, 1]) * np.sqrt(beta / 3.8), "norm").pvalue > 1e-3


def test_diagnostics_use_the_base_temperature_only():
    """PTSampler.diagnostics(): mean / variance of the beta = 1 chains (one per ladder), not of all rungs pooled.
    benchmark_gauss2d_corr has unit marginal variances; the tempered rungs have 1 / beta."""
    from riemann_b200 import PTSampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    pt = PTSampler(benchmarks.benchmark_gauss2d_corr, MetropolisRandomWalk(0.5 * np.eye(2)), np.ones(2), K=2048, seed=4)
    pt.run(500, trace=False)
    pt._sampler.reset_diagnostics()
    pt.run(3000, trace=False)
    dg = pt.diagnostics()
    flat = pt._sampler.diagnostics(allreduce=False)
    assert dg["chains"] == 2048 and dg["temperatures"] == 5
    assert np.all(np.abs(dg["var"][:2] - 1.0) < 0.06), dg["var"]
    assert np.all(np.abs(dg["mean"][:2]) < 0.05), dg["mean"]
    assert np.all(flat["var"][:2] > 3.0)                     # pooled over beta = 1 .. 1/16: (1+2+4+8+16)/5 = 6.2
    assert 0.0 < dg["accept_rate_all_rungs"] < 1.0 and dg["swap_fraction"] == 0.1


def test_dense_path_ladders_replay_and_sample(golden):
    """d = 12 (dense path): 7 ladders fed the reference's stream each reproduce it; in Philox mode every rung samples
    its tempered target N(mu, C / beta) and the carried log-posterior is beta * logL of the state."""
    from scipy import stats
    g = golden("pt_rw_gauss12d")
    K, T = 7, 300
    pt = _pt(g, 12, K=K)
    rep = lambda a: np.concatenate([a[:T]] * K, axis=1)
    pt.run_injected(rep(g["usel"]), rep(g["xi"]), rep(g["u"]))
    for rung in (0, 2, 4):
        th, lp = pt.samplers[rung]._chain_thetas, pt.samplers[rung]._chain_logpost
        for l in (0, 3, 6):
            assert relerr(th[:, l], g["thetas"][rung, :T + 1]) < TOL
            assert relerr(lp[:, l], g["logpost"][rung, :T + 1]) < TOL
    pt2 = _pt(g, 12, K=400, seed=2)
    pt2.run(4000, trace=False)
    th = pt2._thetas[-1]                                                      # [K, nt, d]
    Linv = np.linalg.inv(np.linalg.cholesky(g["C"]))
    for i, beta in enumerate(pt2.betas):
        z = (th[:, i] - g["mu"]) @ Linv.T * np.sqrt(beta)                     # whitened: N(0, I) per coordinate
        assert stats.kstest(z[:, 0], "norm").pvalue > 1e-3, (i, beta)
        assert stats.kstest(z[:, 7], "norm").pvalue > 1e-3, (i, beta)
    m = device_gauss(g, 12)
    ll = m.log_posterior_batch(th.reshape(-1, 12)).cpu().numpy().reshape(400, -1)
    assert relerr(pt2._logpost[-1], ll * pt2.betas[None, :]) < 1e-10
    dg = pt2.diagnostics()
    assert dg["chains"] == 400 and dg["temperatures"] == 5


def test_logistic_ladder_matches_reference_and_samples(golden):
    """Tempering on the logistic model (prior untempered, ptsampler.py:33-37): the reference's PTSampler stream replayed on
    the device; then many Philox ladders whose carried log-posterior is lprior + beta * logL of the state they hold."""
    from riemann_b200 import PTSampler
    from riemann_b200.models.logistic import LogisticRegression
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    from oracle import riemann_port as port
    g = golden("pt_rw_logistic")
    dm = LogisticRegression(g["X"], g["y"], float(g["prior_var"]))
    pt = PTSampler(dm, MetropolisRandomWalk(g["C0"]), g["thetas"][0, 0])
    pt.run_injected(g["usel"], g["xi"], g["u"])
    for i in range(len(pt.betas)):
        assert relerr(np.array(pt.samplers[i]._chain_thetas), g["thetas"][i]) < TOL
        assert relerr(np.array(pt.samplers[i]._chain_logpost), g["logpost"][i]) < TOL
    K = 64
    pt2 = PTSampler(dm, MetropolisRandomWalk(g["C0"]), g["thetas"][0, 0], K=K, seed=6)
    pt2.run(400, trace=False)
    th, lp = pt2._thetas[-1], pt2._logpost[-1]                              # [K, nt, d], [K, nt]
    om = port.LogisticRegression(g["X"], g["y"], float(g["prior_var"]))
    for l in (0, 17, 63):
        for i, beta in enumerate(pt2.betas):
            ref = om.log_prior(th[l, i]) + beta * om.log_likelihood(th[l, i])
            assert abs(lp[l, i] - ref) < 1e-8 * max(1.0, abs(ref))
    assert np.any(th[:, 0] != th[:, 1])
    dg = pt2.diagnostics()
    assert dg["chains"] == K
