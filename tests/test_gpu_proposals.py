"""
GPU parity of the PROPOSALS themselves: the proposed
state and log q(theta'|theta)/q(theta|theta') that Proposal.propose returns
(riemann/proposals/proposal.py:10-17), step by step against the reference's recorded
proposals (fixture `prop_thetas`) and the oracle's logqratio, plus the single-point
`propose()` protocol call and checkpoint/resume.  Tolerance 1e-9 (mMALA 1e-8), fp64 vs fp64.
"""
import numpy as np
import pytest

from gpu_helpers import relerr, device_gauss, oracle_gauss

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _pair(name, g, d):
    """(device model, device proposal, oracle model, oracle proposal) for a Gaussian fixture."""
    from oracle import riemann_port as port
    from riemann_b200.proposals import randomwalk as rw, hamiltonian as hm
    dm, om = device_gauss(g, d), oracle_gauss(g, d)
    M = g["M"] if "M" in g else None
    ns = int(g["nsteps"]) if "nsteps" in g else 1
    if name.startswith("rw_"):
        return dm, rw.MetropolisRandomWalk(g["C0"]), om, port.MetropolisRandomWalk(g["C0"])
    if name.startswith("pcn_"):
        return dm, rw.pCN(g["C0"], float(g["rho"])), om, port.pCN(g["C0"], float(g["rho"]))
    if name.startswith("adaptrw"):
        return dm, rw.AdaptScaleRandomWalk(g["C0"]), om, port.AdaptScaleRandomWalk(g["C0"])
    if name.startswith("adapt"):
        return (dm, hm.AdaptScaleHMC(float(g["eps"]), ns, dm.grad_log_likelihood, M=M),
                om, port.AdaptScaleHMC(float(g["eps"]), ns, om.grad_log_likelihood, M=M))
    return (dm, hm.VanillaHMC(float(g["eps"]), ns, dm.grad_log_likelihood, M=M),
            om, port.VanillaHMC(float(g["eps"]), ns, om.grad_log_likelihood, M=M))


CASES = [("rw_gauss2d", 2), ("rw_gauss5d", 5), ("adaptrw_gauss2d", 2), ("pcn_gauss2d", 2),
         ("mala_gauss5d", 5), ("hmc5_gauss2d", 2), ("hmc3_mass_gauss2d", 2), ("mala_mass_gauss5d", 5),
         ("adapthmc5_gauss2d", 2), ("rw_dense_gauss12d", 12), ("adaptmala_gauss12d", 12),
         ("rw_gauss100d", 100), ("mala_gauss100d", 100), ("mala_gauss1000d", 1000),
         ("hmc4_gauss12d", 12), ("adapthmc3_gauss12d", 12), ("hmc5_gauss100d", 100), ("pcn_gauss12d", 12)]


@pytest.mark.parametrize("name,d", CASES)
def test_every_proposal_and_logqratio_matches_reference(golden, name, d):
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    g = golden(name)
    dm, dp, om, op = _pair(name, g, d)
    T = min(len(g["u"]), 300)
    s = Sampler(dm, dp, g["thetas"][0])
    ex = s.run_injected(xi=g["xi"][:T], u=g["u"][:T])
    assert relerr(ex["prop_theta"][:, 0], g["prop_thetas"][:T]) < TOL       # the reference's own proposals
    o = port.Sampler(om, op, g["thetas"][0], draws=port.VectorTapeDraws(g["xi"], g["u"]))
    lqr = np.empty(T)
    for t in range(T):
        o.sample()
        lqr[t] = o.last_proposal[2]
    scale = max(1.0, np.max(np.abs(lqr)))
    assert np.max(np.abs(ex["logqratio"][:, 0] - lqr)) < 1e-8 * scale


def test_single_point_propose_protocol(golden):
    """Proposal.propose(theta) -> (theta', logqratio), drawing from numpy's global stream like
    the reference: under the fixture's seed the first call reproduces the reference's first proposal."""
    from oracle import riemann_port as port
    for name, d in [("rw_gauss2d", 2), ("mala_gauss5d", 5), ("pcn_gauss2d", 2), ("mala_gauss100d", 100)]:
        g = golden(name)
        dm, dp, om, op = _pair(name, g, d)
        np.random.seed(int(g["seed"]))
        thp, lqr = dp.propose(g["thetas"][0])
        assert relerr(thp, g["prop_thetas"][0]) < TOL
        op.draws = port.VectorTapeDraws(g["xi"][:1], g["u"][:1])
        _, lqr_o = op.propose(g["thetas"][0])
        assert abs(lqr - lqr_o) < 1e-9 * max(1.0, abs(lqr_o))


def test_changepoint_propose_protocol_and_logqratio(golden):
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams, ChangepointRegression1D
    from riemann_b200.proposals.changepoint import ChangepointRegression1DProp
    g = golden("changepoint")
    xmin, xmax, lamb, kmax, alpha, beta, hscale = g["hyper"]
    dm = ChangepointRegression1D(g["x"], g["y"], xmin, xmax, lamb, int(kmax), alpha, beta)
    dp = ChangepointRegression1DProp(dm, hscale)
    om = port.ChangepointRegression1D(g["x"], g["y"], xmin, xmax, lamb, kmax, alpha, beta)
    op = port.ChangepointRegression1DProp(om, hscale)
    # (1) every proposal of a replayed chain equals the oracle's proposal, incl. birth/death log|J|
    T = 1500
    th0 = ChangepointParams([2.0], [1.0, 3.0], 0.1)
    s = Sampler(dm, dp, th0)
    ex = s.run_injected(tape=g["tape"][2, :T])
    o = port.Sampler(om, op, port.ChangepointParams([2.0], [1.0, 3.0], 0.1), draws=port.SlotTapeDraws(g["tape"][2]))
    ndim = 0
    with np.errstate(all="ignore"):
        for t in range(T):
            o.sample()
            thp, _, lqr, _ = o.last_proposal
            k = len(thp.cpx)
            assert ex["prop_k"][t, 0] == k
            assert relerr(ex["prop_cpx"][t, 0, :k], thp.cpx) < TOL and relerr(ex["prop_cpv"][t, 0, :k + 1], thp.cpv) < TOL
            assert abs(ex["prop_sig"][t, 0] - thp.sig) < TOL
            assert abs(ex["logqratio"][t, 0] - lqr) < 1e-9
            ndim += (lqr != 0.0)
    assert ndim > 100                                   # trans-dimensional moves were exercised
    # (2) the protocol call draws from numpy's stream in the reference's order
    for seed in range(5):
        th = port.ChangepointParams(g["cpx"][0, 700, :g["k"][0, 700]], g["cpv"][0, 700, :g["k"][0, 700] + 1], g["sig"][0, 700])
        np.random.seed(seed)
        op.draws = port.LiveDraws()
        want, lqr_o = op.propose(th)
        np.random.seed(seed)
        got, lqr_d = dp.propose(ChangepointParams(th.cpx, th.cpv, th.sig))
        assert len(got.cpx) == len(want.cpx)
        assert relerr(got.cpx, want.cpx) < TOL and relerr(got.cpv, want.cpv) < TOL
        assert abs(float(got.sig) - want.sig) < TOL and abs(lqr_d - lqr_o) < 1e-9


@pytest.mark.parametrize("name,tol", [("mala_logistic", 1e-9), ("mmala_logistic", 1e-8)])
def test_logistic_proposals_and_logqratio(golden, name, tol):
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.models.logistic import LogisticRegression
    from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA
    g = golden(name)
    dm = LogisticRegression(g["X"], g["y"], float(g["prior_var"]))
    om = port.LogisticRegression(g["X"], g["y"], float(g["prior_var"]))
    eps = float(g["eps"])
    dp, op = ((MALA(eps, dm.grad_log_posterior), port.MALA(eps, om.grad_log_posterior)) if name == "mala_logistic"
              else (SimplifiedMMALA(eps, dm), port.SimplifiedMMALA(eps, om)))
    T = 200
    s = Sampler(dm, dp, g["thetas"][0])
    ex = s.run_injected(xi=g["xi"][:T], u=g["u"][:T])
    assert relerr(ex["prop_theta"][:, 0], g["prop_thetas"][:T]) < tol
    o = port.Sampler(om, op, g["thetas"][0], draws=port.VectorTapeDraws(g["xi"], g["u"]))
    for t in range(T):
        o.sample()
        assert abs(ex["logqratio"][t, 0] - o.last_proposal[2]) < 1e-7 * max(1.0, abs(o.last_proposal[2]))


def test_checkpoint_resume_is_bit_exact():
    """(state, Philox step counter, adapt state) is the complete resumable state."""
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import AdaptScaleRandomWalk
    m = benchmarks.benchmark_gauss2d_corr
    a = Sampler(m, AdaptScaleRandomWalk(0.01 * np.eye(2)), np.ones(2), K=64, seed=4)
    a.run(200, trace=False)
    ck = a.get_checkpoint()
    a.run(300, trace=False)
    b = Sampler(m, AdaptScaleRandomWalk(0.01 * np.eye(2)), np.zeros(2), K=64, seed=4)
    b.set_checkpoint(ck)
    b.run(300, trace=False)
    assert np.array_equal(np.asarray(a._chain_thetas[-1]), np.asarray(b._chain_thetas[-1]))
    assert np.array_equal(a.proposal.scale, b.proposal.scale)


def test_checkpoint_is_refused_where_the_state_cannot_be_saved():
    """The covariance-adapting proposals and AdaptScalepCN keep per-chain state the ABI cannot read back (Haario
    accumulators, compounding rho): get_checkpoint must refuse instead of resuming silently different chains."""
    from riemann_b200 import ParameterError, Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import AdaptCovRandomWalk, AdaptScalepCN
    m = benchmarks.benchmark_gauss2d_corr
    for prop in (AdaptCovRandomWalk(0.1 * np.eye(2)), AdaptScalepCN(np.eye(2), 0.5)):
        s = Sampler(m, prop, np.ones(2), K=8, seed=1)
        s.run(20, trace=False)
        with pytest.raises(ParameterError):
            s.get_checkpoint()
