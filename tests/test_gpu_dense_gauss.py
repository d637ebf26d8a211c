"""
GPU parity for the dense Gaussian path (riemann_b200/csrc/dense.cu, d > 8): one fp64
tensor-core GEMM  V' = Y' P  per MH step with the quadratic form / log q ratio fused into
its epilogue.  The stream the REFERENCE drew (VanillaHMC(eps, 1, grad) = MALA and
MetropolisRandomWalk on MultiGaussianDist, d = 12 / 100 / 1000) is replayed.

Tolerance: 1e-9 relative-or-absolute on states and log-posteriors (fp64 vs fp64; the
device multiplies by P = C^-1 where numpy solves with the Cholesky factor; at d = 1000 the
condition number 9e3 of 0.1 I + 0.9 11^T amplifies round-off to ~1e-11); decisions identical.
"""
import numpy as np
import pytest

from gpu_helpers import relerr, device_gauss, oracle_gauss

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _prop(name, g, m):
    from riemann_b200.proposals import randomwalk as rw, hamiltonian as hm
    if name.startswith("rw_"):
        return rw.MetropolisRandomWalk(g["C0"])
    if name.startswith("adaptrw_"):
        return rw.AdaptScaleRandomWalk(g["C0"])
    if name.startswith("pcn"):
        return rw.pCN(g["C0"], float(g["rho"]))
    M = g["M"] if "M" in g else None
    if name.startswith("adaptmalamass"):
        return hm.AdaptScaleHMC(float(g["eps"]), 1, m.grad_log_likelihood, M=M)
    if name.startswith("hmcmass"):
        return hm.VanillaHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_likelihood, M=M)
    if name.startswith("adaptmala"):
        return hm.AdaptScaleHMC(float(g["eps"]), 1, m.grad_log_likelihood)
    if name.startswith("adapthmc"):
        return hm.AdaptScaleHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_likelihood)
    if name.startswith("hmc"):
        return hm.VanillaHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_likelihood)
    return hm.MALA(float(g["eps"]), m.grad_log_likelihood)


@pytest.mark.parametrize("name,d", [("rw_gauss100d", 100), ("rw_dense_gauss12d", 12),
                                    ("adaptrw_gauss12d", 12), ("adaptmala_gauss12d", 12),
                                    ("mala_gauss100d", 100), ("mala_gauss1000d", 1000),
                                    # "next" row N1: leapfrog with Nsteps > 1 on the dense path
                                    ("hmc4_gauss12d", 12), ("adapthmc3_gauss12d", 12), ("hmc5_gauss100d", 100),
                                    # leapfrog with a mass matrix (hamiltonian.py:18-21) on the dense path
                                    ("hmcmass3_gauss12d", 12), ("adaptmalamass_gauss12d", 12),
                                    # "next" row N2: pCN on the dense path
                                    ("pcn_gauss12d", 12)])
def test_injected_chain_matches_reference(golden, name, d):
    from riemann_b200 import Sampler
    g = golden(name)
    m = device_gauss(g, d)
    p = _prop(name, g, m)
    s = Sampler(m, p, g["thetas"][0])
    ex = s.run_injected(xi=g["xi"], u=g["u"])
    th = np.array(s._chain_thetas)
    assert relerr(th, g["thetas"]) < TOL
    assert relerr(s._chain_logpost, g["logpost"]) < TOL
    assert relerr(ex["prop_logpost"][:, 0], g["prop_logpost"]) < TOL
    assert np.array_equal(ex["accepted"][:, 0], np.any(g["thetas"][1:] != g["thetas"][:-1], axis=1))
    if "scales" in g:
        assert abs(p.scale - g["scales"][-1]) < 1e-11 * g["scales"][-1]
        assert abs(p.accept_rate - float(g["accept_rate"])) < 1e-15


def test_ragged_chain_count_and_per_chain_streams(golden):
    """K = 131 chains (not a multiple of the 128-row GEMM tile), each replaying a different
    window of the reference stream from a different start state; checked against the oracle."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import MALA
    g = golden("mala_gauss100d")
    K, T, d = 131, 40, 100
    rng = np.random.default_rng(5)
    starts = rng.integers(0, len(g["u"]) - T, K)
    xi = np.stack([g["xi"][s0:s0 + T] for s0 in starts], axis=1)
    u = np.stack([g["u"][s0:s0 + T] for s0 in starts], axis=1)
    th0 = rng.standard_normal((K, d)) * 0.5
    m = device_gauss(g, d)
    s = Sampler(m, MALA(float(g["eps"]), m.grad_log_likelihood), th0)
    s.run_injected(xi=xi, u=u)
    om = oracle_gauss(g, d)
    for c in (0, 1, 64, 127, 128, 130):
        o = port.Sampler(om, port.MALA(float(g["eps"]), om.grad_log_likelihood), th0[c],
                         draws=port.VectorTapeDraws(xi[:, c], u[:, c]))
        o.run(T)
        assert relerr(s._chain_thetas[:, c], np.array(o._chain_thetas)) < TOL
        assert relerr(s._chain_logpost[:, c], np.array(o._chain_logpost)) < TOL


def test_ragged_mass_matrix_trajectories_match_oracle():
    """HMC with a mass matrix on the dense path at K = 131 (not a multiple of the 128-row tile) and d = 37
    (padded to 48): every chain has its own start and noise; six chains are replayed by the oracle."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.models.gaussian import MultiGaussianDist
    from riemann_b200.proposals.hamiltonian import VanillaHMC
    K, T, d = 131, 12, 37
    rng = np.random.default_rng(11)
    A = rng.standard_normal((d, d))
    C = A @ A.T / d + 0.3 * np.eye(d)
    mu = rng.standard_normal(d)
    M = np.linalg.inv(C) + np.diag(rng.uniform(0.1, 0.5, d))
    M = 0.5 * (M + M.T)
    xi = rng.standard_normal((T, K, d))
    u = rng.uniform(size=(T, K))
    th0 = mu + 0.5 * rng.standard_normal((K, d))
    m = MultiGaussianDist(mu, C)
    s = Sampler(m, VanillaHMC(0.45, 3, m.grad_log_likelihood, M=M), th0)
    s.run_injected(xi=xi, u=u)
    om = port.MultiGaussianDist(mu, C)
    nacc = 0
    for c in (0, 1, 64, 127, 128, 130):
        o = port.Sampler(om, port.VanillaHMC(0.45, 3, om.grad_log_likelihood, M=M), th0[c],
                         draws=port.VectorTapeDraws(xi[:, c], u[:, c]))
        o.run(T)
        ot = np.array(o._chain_thetas)
        assert relerr(s._chain_thetas[:, c], ot) < TOL
        assert relerr(s._chain_logpost[:, c], np.array(o._chain_logpost)) < TOL
        nacc += int(np.any(ot[1:] != ot[:-1], axis=1).sum())
    assert 0 < nacc < 6 * T          # both accepts and rejects occur in the replayed chains


def test_thinned_trace_and_resume(golden):
    from riemann_b200 import Sampler
    from riemann_b200.proposals.hamiltonian import MALA
    g = golden("mala_gauss100d")
    m = device_gauss(g, 100)
    s = Sampler(m, MALA(float(g["eps"]), m.grad_log_likelihood), g["thetas"][0])
    s.run(100, 20, 7, inject={"xi": g["xi"][:100, None, :], "u": g["u"][:100, None]})
    idx = list(range(20, 101, 7))
    assert relerr(np.array(s._chain_thetas), g["thetas"][idx]) < TOL
    s.run(50, inject={"xi": g["xi"][100:150, None, :], "u": g["u"][100:150, None]})     # resumes
    assert relerr(np.array(s._chain_thetas), g["thetas"][100:151]) < TOL


def test_philox_mala_d100_matches_target():
    """Distributional gate: benchmark_gauss100d_corr (benchmarks.py:25-26) has mean 0, unit
    marginal variances and pairwise correlation 0.9."""
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.hamiltonian import MALA
    m = benchmarks.benchmark_gauss100d_corr
    K = 4096
    rng = np.random.default_rng(0)
    th0 = rng.standard_normal((K, 100)) * np.sqrt(0.1) + rng.standard_normal((K, 1)) * np.sqrt(0.9)
    s = Sampler(m, MALA(0.12, m.grad_log_likelihood), th0, seed=3)
    s.run(400, trace=False)
    s.reset_diagnostics()
    s.run(400, trace=False)
    dg = s.diagnostics(allreduce=False)
    th = np.asarray(s._chain_thetas[-1])
    assert 0.85 < dg["accept_rate"] < 0.99
    assert np.all(np.abs(th.mean(0)) < 0.08)
    assert np.all(np.abs(th.var(0) - 1.0) < 0.1)
    resid = th - th.mean(1, keepdims=True)
    assert abs(resid.var() - 0.1 * 99 / 100) < 0.005           # the 99 stiff directions: variance 0.1
    lp = np.asarray(s._chain_logpost[-1])
    want = m.log_posterior_batch(th).cpu().numpy()
    assert relerr(lp, want) < 1e-10                             # carried log-posterior == fresh evaluation


def test_full_size_config3_self_consistency():
    """BASELINE config 3 shape: d = 1000, 16,384 chains, MALA.  Property checks that do not
    need the O(d^3) CPU oracle: the carried log-posterior equals a fresh evaluation, and the
    energy error of a MALA step is O(eps^3) so acceptance is high."""
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.hamiltonian import MALA
    m = benchmarks.gauss_corr(1000)
    K = 16384
    rng = np.random.default_rng(1)
    th0 = rng.standard_normal((K, 1000)) * np.sqrt(0.1) + rng.standard_normal((K, 1)) * np.sqrt(0.9)
    s = Sampler(m, MALA(0.08, m.grad_log_likelihood), th0, seed=9)
    s.run(10, trace=False)
    dg = s.diagnostics(allreduce=False)
    assert 0.5 < dg["accept_rate"] <= 1.0
    th, lp = s.state_tensors()
    want = m.log_posterior_batch(th[:512])
    assert relerr(lp[:512].cpu().numpy(), want.cpu().numpy()) < 1e-10


def test_philox_hmc5_d100_matches_target():
    """N1 distributional gate: 5 leapfrog steps per proposal on benchmark_gauss100d_corr."""
    from scipy import stats
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.hamiltonian import VanillaHMC
    m = benchmarks.benchmark_gauss100d_corr
    K = 4096
    rng = np.random.default_rng(0)
    th0 = rng.standard_normal((K, 100)) * np.sqrt(0.1) + rng.standard_normal((K, 1)) * np.sqrt(0.9)
    s = Sampler(m, VanillaHMC(0.1, 5, m.grad_log_likelihood), th0, seed=4)
    s.run(300, trace=False)
    dg = s.diagnostics(allreduce=False)
    th = np.asarray(s._chain_thetas[-1])
    assert 0.8 < dg["accept_rate"] <= 1.0
    assert stats.kstest(th[:, 3], "norm").pvalue > 1e-3
    assert stats.kstest(th.mean(1) / np.sqrt(0.9 + 0.1 / 100), "norm").pvalue > 1e-3
    assert stats.kstest((th[:, 0] - th[:, 1]) / np.sqrt(0.2), "norm").pvalue > 1e-3
    lp = np.asarray(s._chain_logpost[-1])
    assert relerr(lp, m.log_posterior_batch(th).cpu().numpy()) < 1e-10


@pytest.mark.parametrize("precision", ["f64", "tf32x3"])
def test_diagnostics_are_those_of_theta_when_mu_is_not_zero(precision):
    """The dense paths keep the centred state y = theta - mu; the tracked functionals (first 7 coordinates and
    mean(theta)) must still be those of theta (riemann/models/gaussian.py:49-52 is a density of theta)."""
    from riemann_b200 import Sampler
    from riemann_b200.models.gaussian import MultiGaussianDist
    from riemann_b200.proposals.hamiltonian import MALA
    d, K = 24, 4096
    rng = np.random.default_rng(2)
    mu = 3.0 + rng.standard_normal(d)
    C = 0.5 * np.eye(d) + 0.5 * np.ones((d, d)) / d
    m = MultiGaussianDist(mu, C)
    th0 = mu[None] + rng.multivariate_normal(np.zeros(d), C, size=K)
    s = Sampler(m, MALA(0.3, m.grad_log_likelihood), th0, seed=3, precision=precision)
    s.run(50, trace=False)
    s.reset_diagnostics()
    s.run(100, trace=False)
    dg = s.diagnostics(allreduce=False)
    assert np.all(np.abs(dg["mean"][:7] - mu[:7]) < 0.05)
    assert abs(dg["mean"][7] - mu.mean()) < 0.03
