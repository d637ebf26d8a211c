"""
CPU: the numpy port (oracle/riemann_port.py) replays the stream the REFERENCE drew
(tests/golden/*.npz, produced by oracle/gen_golden.py from /root/reference) and must
reproduce the reference's chains.  This is what pins the oracle.

Tolerance: fp64 round-off only (the port uses triangular solves where the reference
uses a general LU solve on the same triangular factor): 1e-9 absolute on states and
log-posteriors; accept/reject decisions identical.
"""
import numpy as np
import pytest

from oracle import riemann_port as port

TOL = 1e-9


def _err(a, b):
    """max |a-b| / max(1, |b|): absolute near zero, relative for huge values."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)), initial=0.0)


def _gauss(g, d):
    if "C" in g:
        return port.MultiGaussianDist(g["mu"], g["C"])
    return port.benchmark_gauss(d, corr=(d > 1))


def _replay(g, model, proposal, theta0):
    T = len(g["u"])
    s = port.Sampler(model, proposal, theta0, draws=port.VectorTapeDraws(g["xi"], g["u"]))
    prop_lp = np.empty(T)
    for t in range(T):
        s.sample()
        prop_lp[t] = s.last_proposal[1]
    thetas = np.array([np.atleast_1d(x) for x in s._chain_thetas])
    lps = np.array(s._chain_logpost)
    assert _err(thetas, g["thetas"]) < TOL
    assert _err(lps, g["logpost"]) < TOL
    assert _err(prop_lp, g["prop_logpost"]) < TOL
    # decisions identical
    assert np.array_equal(np.any(thetas[1:] != thetas[:-1], axis=1),
                          np.any(g["thetas"][1:] != g["thetas"][:-1], axis=1))
    return s


@pytest.mark.parametrize("name,d", [("rw_gauss1d", 1), ("rw_gauss2d", 2), ("rw_gauss5d", 5),
                                    ("rw_gauss100d", 100), ("rw_dense_gauss12d", 12)])
def test_rw(golden, name, d):
    g = golden(name)
    _replay(g, _gauss(g, d), port.MetropolisRandomWalk(g["C0"]), g["thetas"][0])


def test_adapt_scale_rw(golden):
    g = golden("adaptrw_gauss2d")
    prop = port.AdaptScaleRandomWalk(g["C0"])
    scales = [prop.scale]
    orig = prop.adapt
    prop.adapt = lambda th: (orig(th), scales.append(prop.scale))
    _replay(g, _gauss(g, 2), prop, g["thetas"][0])
    assert np.allclose(scales, g["scales"], rtol=1e-12, atol=0)
    assert abs(prop.accept_rate - float(g["accept_rate"])) < 1e-15


@pytest.mark.parametrize("name,d", [("mala_gauss2d", 2), ("mala_gauss5d", 5),
                                    ("mala_gauss100d", 100), ("mala_gauss1000d", 1000)])
def test_mala_is_hmc1(golden, name, d):
    g = golden(name)
    m = _gauss(g, d)
    _replay(g, m, port.MALA(float(g["eps"]), m.grad_log_likelihood), g["thetas"][0])


@pytest.mark.parametrize("name,d", [("hmc5_gauss2d", 2), ("hmc3_mass_gauss2d", 2), ("hmc4_gauss12d", 12),
                                    ("hmc5_gauss100d", 100), ("hmcmass3_gauss12d", 12)])
def test_hmc(golden, name, d):
    g = golden(name)
    m = _gauss(g, d)
    M = g["M"] if "M" in g else None
    _replay(g, m, port.VanillaHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_likelihood, M=M),
            g["thetas"][0])


def test_mala_mass(golden):
    g = golden("mala_mass_gauss5d")
    m = _gauss(g, 5)
    _replay(g, m, port.VanillaHMC(float(g["eps"]), 1, m.grad_log_likelihood, M=g["M"]),
            g["thetas"][0])


@pytest.mark.parametrize("name,d", [("adapthmc5_gauss2d", 2), ("adapthmc3_gauss12d", 12),
                                    ("adaptmalamass_gauss12d", 12)])
def test_adapt_scale_hmc(golden, name, d):
    g = golden(name)
    m = _gauss(g, d)
    prop = port.AdaptScaleHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_likelihood,
                              M=g["M"] if "M" in g else None)
    _replay(g, m, prop, g["thetas"][0])
    assert abs(prop.scale - g["scales"][-1]) < 1e-12 * g["scales"][-1]


@pytest.mark.parametrize("name", ["adaptrw_gauss12d", "adaptmala_gauss12d"])
def test_adaptive_d12(golden, name):
    g = golden(name)
    m = _gauss(g, 12)
    prop = (port.AdaptScaleRandomWalk(g["C0"]) if name.startswith("adaptrw")
            else port.AdaptScaleHMC(float(g["eps"]), 1, m.grad_log_likelihood))
    _replay(g, m, prop, g["thetas"][0])
    assert abs(prop.scale - g["scales"][-1]) < 1e-12 * g["scales"][-1]


@pytest.mark.parametrize("name,d", [("pcn_gauss2d", 2), ("pcn_gauss12d", 12)])
def test_pcn(golden, name, d):
    g = golden(name)
    _replay(g, _gauss(g, d), port.pCN(g["C0"], float(g["rho"])), g["thetas"][0])


@pytest.mark.parametrize("name", ["mala_logistic", "mmala_logistic", "hmc3_logistic", "adapthmc4_logistic",
                                  "hmcmass3_logistic", "adaptmalamass_logistic"])
def test_logistic(golden, name):
    """Port model (+ port mMALA) were driven through the REFERENCE Sampler/VanillaHMC/AdaptScaleHMC."""
    g = golden(name)
    m = port.LogisticRegression(g["X"], g["y"], float(g["prior_var"]))
    if name == "mala_logistic":
        prop = port.MALA(float(g["eps"]), m.grad_log_posterior)
    elif name == "hmc3_logistic":
        prop = port.VanillaHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_posterior)
    elif name == "adapthmc4_logistic":
        prop = port.AdaptScaleHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_posterior)
    elif name == "hmcmass3_logistic":           # fixed dense mass matrix, hamiltonian.py:70-89
        prop = port.VanillaHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_posterior, M=g["M"])
    elif name == "adaptmalamass_logistic":
        prop = port.AdaptScaleHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_posterior, M=g["M"])
    else:
        prop = port.SimplifiedMMALA(float(g["eps"]), m)
    _replay(g, m, prop, g["thetas"][0])


def test_changepoint(golden):
    g = golden("changepoint")
    xmin, xmax, lamb, kmax, alpha, beta, hscale = g["hyper"]
    nch, T = g["tape"].shape[:2]
    for c in range(nch):
        model = port.ChangepointRegression1D(g["x"], g["y"], xmin, xmax, lamb, kmax, alpha, beta)
        prop = port.ChangepointRegression1DProp(model, hscale)
        k0 = int(g["k"][c, 0])
        th0 = port.ChangepointParams(g["cpx"][c, 0, :k0], g["cpv"][c, 0, :k0 + 1], g["sig"][c, 0])
        s = port.Sampler(model, prop, th0, draws=port.SlotTapeDraws(g["tape"][c]))
        with np.errstate(all="ignore"):
            for t in range(T):
                s.sample()
                assert abs(s.last_proposal[1] - g["prop_logpost"][c, t]) < TOL or \
                    (np.isinf(s.last_proposal[1]) and np.isinf(g["prop_logpost"][c, t]))
        for t in range(0, T + 1):
            th = s._chain_thetas[t]
            k = len(th.cpx)
            assert k == g["k"][c, t]
            assert np.max(np.abs(th.cpx - g["cpx"][c, t, :k]), initial=0) < TOL
            assert np.max(np.abs(th.cpv - g["cpv"][c, t, :k + 1])) < TOL
            assert abs(th.sig - g["sig"][c, t]) < TOL
        assert np.max(np.abs(np.array(s._chain_logpost) - g["logpost"][c])) < TOL


def test_live_stream_matches_tape(golden):
    """Under np.random.seed the port consumes numpy's stream in the reference's order."""
    g = golden("rw_gauss2d")
    np.random.seed(int(g["seed"]))
    s = port.Sampler(port.benchmark_gauss(2), port.MetropolisRandomWalk(g["C0"]), np.ones(2))
    s.run(200)
    assert np.max(np.abs(np.array(s._chain_thetas) - g["thetas"][:201])) < TOL

    g = golden("changepoint")
    xmin, xmax, lamb, kmax, alpha, beta, hscale = g["hyper"]
    model = port.ChangepointRegression1D(g["x"], g["y"], xmin, xmax, lamb, kmax, alpha, beta)
    th0 = port.ChangepointParams(g["cpx"][0, 0, :1], g["cpv"][0, 0, :2], g["sig"][0, 0])
    np.random.seed(int(g["seed0"]))
    s = port.Sampler(model, port.ChangepointRegression1DProp(model, hscale), th0)
    with np.errstate(all="ignore"):
        s.run(1500)
    assert np.max(np.abs(np.array(s._chain_logpost) - g["logpost"][0, :1501])) < TOL


@pytest.mark.parametrize("name,d", [("pt_rw_gauss2d", 2), ("pt_rw_gauss5d", 5), ("pt_rw_gauss12d", 12),
                                    ("pt_rw_logistic", 6)])
def test_parallel_tempering(golden, name, d):
    """N3: port.PTSampler replays the recorded stream of the reference's PTSampler (ptsampler.py:41-127):
    every temperature's chain, log-posterior and swap decision."""
    g = golden(name)
    m = port.LogisticRegression(g["X"], g["y"], float(g["prior_var"])) if name == "pt_rw_logistic" else _gauss(g, d)
    pt = port.PTSampler(m, port.MetropolisRandomWalk(g["C0"]), g["thetas"][0, 0],
                        draws=port.PTTapeDraws(g["usel"], g["xi"], g["u"]))
    assert np.array_equal(pt.betas, g["betas"]) and pt.Pswap == float(g["pswap"])
    T = g["usel"].shape[0]
    with np.errstate(all="ignore"):
        pt.run(T)
    for i, s in enumerate(pt.samplers):
        assert np.max(np.abs(np.array(s._chain_thetas) - g["thetas"][i])) < 1e-12
        assert np.max(np.abs(np.array(s._chain_logpost) - g["logpost"][i])) < 1e-9
    assert pt._chain_thetas is pt.samplers[0]._chain_thetas


@pytest.mark.parametrize("name,d", [("adaptcov_gauss2d", 2), ("adaptcov_smooth_gauss2d", 2),
                                    ("adaptcov_gauss5d", 5), ("adaptcov_marg_gauss5d", 5),
                                    ("adaptcov_smooth_gauss5d", 5), ("adaptcov_smooth_marg_gauss5d", 5)])
def test_adapt_cov_random_walk(golden, name, d):
    """N5: Haario-style covariance adaptation (adaptive.py:38-103 through randomwalk.py:40-56)."""
    g = golden(name)
    prop = port.AdaptCovRandomWalk(g["C0"], t_adapt=float(g["t_adapt"]), marginalize=bool(g["marginalize"]),
                                   smooth_adapt=bool(g["smooth_adapt"]))
    _replay(g, _gauss(g, d), prop, g["thetas"][0])
    assert np.max(np.abs(prop.L - g["L_final"])) < 1e-12 * np.max(np.abs(g["L_final"]))


def test_example_script_proposals(golden):
    """The remaining proposals of examples/test_randomwalk.py:23-37 (BASELINE config 0 runs AdaptScaleCovHMC)."""
    g = golden("adaptscalecovhmc5_gauss2d")
    m = _gauss(g, 2)
    p = port.AdaptScaleCovHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_likelihood, g["M0"], t_adapt=100,
                              smooth_adapt=True)
    _replay(g, m, p, g["thetas"][0])
    assert abs(p.scale - g["scales"][-1]) < 1e-12 * g["scales"][-1]
    g = golden("adaptscalecovhmc3_mass_gauss2d")
    p = port.AdaptScaleCovHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_likelihood, g["M0"])
    _replay(g, m, p, g["thetas"][0])
    g = golden("adaptscalecov_rw_gauss2d")
    p = port.AdaptScaleCovRandomWalk(g["C0"], t_adapt=float(g["t_adapt"]), smooth_adapt=True)
    _replay(g, m, p, g["thetas"][0])
    assert abs(p.scale - g["scales"][-1]) < 1e-12 * g["scales"][-1]
    g = golden("adaptscalepcn_gauss2d")
    p = port.AdaptScalepCN(g["C0"], float(g["rho"]))
    _replay(g, m, p, g["thetas"][0])
    assert abs(p.scale - g["scales"][-1]) < 1e-12 * g["scales"][-1]
    g = golden("adaptcovhmc5_gauss2d")
    p = port.AdaptCovHMC(float(g["eps"]), int(g["nsteps"]), m.grad_log_likelihood, g["M0"], t_adapt=float(g["t_adapt"]),
                         smooth_adapt=True)
    _replay(g, m, p, g["thetas"][0])
