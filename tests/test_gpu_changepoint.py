"""
GPU parity for the changepoint kernel (riemann_b200/csrc/changepoint.cu): config 2.

* the stream the REFERENCE drew (tests/golden/changepoint.npz: 4 chains x 3000 steps of
  examples/test_changepoint.py's proposal on riemann/models/changepoint.py) is replayed on
  the device; states, log-posteriors and decisions must match at every step;
* pointwise log-posterior / likelihood / prior vs the numpy oracle, including the invalid
  states the reference rejects through nan/inf -> -inf;
* long Philox runs agree distributionally with oracle chains;
* full-size (65,536 chains) self-consistency.

Tolerance: 1e-9 relative-or-absolute (fp64 device vs fp64 numpy; the device sums residuals
per segment from prefix sums, numpy sums them per datum); decisions identical.
"""
import numpy as np
import pytest

from gpu_helpers import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _setup(g=None):
    from oracle import riemann_port as port
    from riemann_b200.models.changepoint import ChangepointRegression1D
    from riemann_b200.proposals.changepoint import ChangepointRegression1DProp
    if g is not None:
        x, y = g["x"], g["y"]
        xmin, xmax, lamb, kmax, alpha, beta, hscale = g["hyper"]
    else:
        pm, pp, _, _ = port.make_changepoint_problem()
        x, y, xmin, xmax, lamb, kmax, alpha, beta, hscale = (pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb,
                                                             pm.kmax, pm.alpha, pm.beta, pp.hscale)
    dm = ChangepointRegression1D(x, y, xmin, xmax, lamb, int(kmax), alpha, beta)
    dp = ChangepointRegression1DProp(dm, hscale)
    om = port.ChangepointRegression1D(x, y, xmin, xmax, lamb, kmax, alpha, beta)
    op = port.ChangepointRegression1DProp(om, hscale)
    return dm, dp, om, op


def test_injected_chains_match_reference(golden):
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    g = golden("changepoint")
    dm, dp, _, _ = _setup(g)
    nch, T = g["tape"].shape[:2]
    th0 = [ChangepointParams(g["cpx"][c, 0, :g["k"][c, 0]], g["cpv"][c, 0, :g["k"][c, 0] + 1],
                             g["sig"][c, 0]) for c in range(nch)]
    s = Sampler(dm, dp, th0)
    ex = s.run_injected(tape=np.transpose(g["tape"], (1, 0, 2)))
    tr = s._chain_thetas
    assert tr.k.shape == (T + 1, nch)
    assert np.array_equal(tr.k.T, g["k"])
    assert relerr(np.transpose(tr.cpx, (1, 0, 2)), g["cpx"]) < TOL
    assert relerr(np.transpose(tr.cpv, (1, 0, 2)), g["cpv"]) < TOL
    assert relerr(tr.sig.T, g["sig"]) < TOL
    assert relerr(s._chain_logpost.T, g["logpost"]) < TOL
    assert relerr(ex["prop_logpost"].T, g["prop_logpost"]) < TOL
    ref_acc = g["logpost"][:, 1:] != g["logpost"][:, :-1]
    # a k=0 cpx move re-proposes the same state (accepted, unchanged): compare where it matters
    moved = ex["accepted"].T & (ex["prop_logpost"].T != g["logpost"][:, :-1])
    assert np.array_equal(moved, ref_acc)


def test_single_chain_api_matches_reference_script(golden):
    """K = 1 keeps the reference's list-of-ChangepointParams history (test_changepoint.py:87)."""
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    g = golden("changepoint")
    dm, dp, _, _ = _setup(g)
    th0 = ChangepointParams([2.0], [1.0, 3.0], 0.1)
    s = Sampler(dm, dp, th0)
    T = 500
    s.run_injected(tape=g["tape"][1, :T])
    assert isinstance(s._chain_thetas, list) and len(s._chain_thetas) == T + 1
    for t in (0, 1, 77, T):
        th = s._chain_thetas[t]
        k = g["k"][1, t]
        assert len(th.cpx) == k and len(th.cpv) == k + 1
        assert relerr(th.cpx, g["cpx"][1, t, :k]) < TOL and relerr(th.cpv, g["cpv"][1, t, :k + 1]) < TOL
    assert relerr(s._chain_logpost, g["logpost"][1, :T + 1]) < TOL


def test_pointwise_including_invalid_states(golden):
    from oracle import riemann_port as port
    from riemann_b200.models.changepoint import ChangepointParams
    g = golden("changepoint")
    dm, _, om, _ = _setup(g)
    rng = np.random.default_rng(0)
    states = []
    for _ in range(200):
        k = int(rng.integers(0, 12))
        states.append((np.sort(rng.uniform(1.0, 3.0, k)), rng.uniform(0.5, 3.5, k + 1), rng.uniform(0.03, 0.5)))
    states += [
        (np.array([2.5, 1.5]), np.array([1.0, 2.0, 3.0]), 0.1),      # unsorted -> -inf
        (np.array([0.5]), np.array([1.0, 2.0]), 0.1),                # below xmin -> -inf
        (np.array([3.5]), np.array([1.0, 2.0]), 0.1),                # above xmax -> -inf
        (np.array([2.0, 2.0]), np.array([1.0, 2.0, 3.0]), 0.1),      # zero gap -> -inf
        (np.array([2.0]), np.array([-1.0, 2.0]), 0.1),               # negative height -> -inf
        (np.array([2.0]), np.array([0.0, 2.0]), 0.1),                # zero height -> -inf
        (np.array([2.0]), np.array([1.0, 2.0]), -0.1),               # sigma < 0 -> -inf
        (np.array([2.0]), np.array([1.0, 2.0]), 0.0),                # sigma = 0 -> -inf
        (np.array([]), np.array([2.0]), 0.2),                        # k = 0
    ]
    dth = [ChangepointParams(*s) for s in states]
    oth = [port.ChangepointParams(*s) for s in states]
    with np.errstate(all="ignore"):
        want = np.array([om.log_posterior(t) for t in oth])
    got = dm.log_posterior_batch(dth)
    assert relerr(got, want) < TOL
    assert np.all(np.isneginf(got[200:208]))
    for i in (0, 5, 208):
        assert abs(dm.log_likelihood(dth[i]) - om.log_likelihood(oth[i])) < 1e-9 * abs(want[i])
        assert abs(dm.log_prior(dth[i]) - om.log_prior(oth[i])) < 1e-9 * max(1, abs(want[i]))


def test_alpha_beta_not_one():
    """gamma(alpha, beta) height prior with alpha != 1 exercises the log(v) branch."""
    from oracle import riemann_port as port
    from riemann_b200.models.changepoint import ChangepointParams, ChangepointRegression1D
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 1, 37))
    y = rng.normal(2.0, 0.3, 37)
    dm = ChangepointRegression1D(x, y, 0.0, 1.0, 3.0, 10, 2.5, 1.7)
    om = port.ChangepointRegression1D(x, y, 0.0, 1.0, 3.0, 10, 2.5, 1.7)
    st = [(np.sort(rng.uniform(0, 1, k)), rng.uniform(0.5, 3, k + 1), 0.3) for k in range(0, 9)]
    got = dm.log_posterior_batch([ChangepointParams(*s) for s in st])
    want = [om.log_posterior(port.ChangepointParams(*s)) for s in st]
    assert relerr(got, want) < TOL


def test_infinite_start_is_always_left(golden):
    """Appendix A.2: Python's min(0, nan) == 0, so a -inf -> anything move is accepted."""
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    from riemann_b200 import _lib
    g = golden("changepoint")
    dm, dp, _, _ = _setup(g)
    s = Sampler(dm, dp, ChangepointParams([2.0], [1.0, 3.0], -0.1))     # sigma < 0: logpost = -inf
    assert s._chain_logpost[0] == -np.inf
    tape = np.zeros((1, _lib.CP_NSLOT))
    tape[0, _lib.CP_SLOT["sel1"]] = 0.1                                  # cpx move
    tape[0, _lib.CP_SLOT["acc"]] = 0.999999
    ex = s.run_injected(tape=tape)
    assert ex["accepted"][0, 0] and s._chain_logpost[-1] == -np.inf


def test_kcap_overflow_is_counted_not_silent(golden):
    from riemann_b200 import Sampler, _lib
    from riemann_b200.models.changepoint import ChangepointParams
    g = golden("changepoint")
    dm, dp, _, _ = _setup(g)
    k = _lib.CP_LANES - 1
    th0 = ChangepointParams(np.linspace(1.1, 2.9, k), np.full(k + 1, 2.0), 0.1)
    s = Sampler(dm, dp, th0)
    tape = np.zeros((3, _lib.CP_NSLOT))
    tape[:, :4] = 0.9                       # -> trans-dimensional, birth
    tape[:, _lib.CP_SLOT["s"]] = 2.03
    tape[:, _lib.CP_SLOT["acc"]] = 1e-12
    s.run_injected(tape=tape)
    dg = s.diagnostics(allreduce=False)
    assert dg["overflows"] == 3 and dg["accept_rate"] == 0.0
    assert len(s._chain_thetas[-1].cpx) == k


def test_philox_run_agrees_with_reference_posterior(golden):
    """Distributional gate vs the REFERENCE sampler: tests/golden/changepoint_posterior.npz holds
    window statistics (MH steps 6000..10000 from the example's start state) of 48 independent
    chains of the unmodified reference, each on numpy's own stream.  4096 device chains over the
    same window must agree within 4 standard errors of the reference estimate."""
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    g = golden("changepoint_posterior")
    dm, dp, _, _ = _setup()
    K = 4096
    s = Sampler(dm, dp, ChangepointParams([2.0], [1.0, 3.0], 0.1), K=K, seed=2024)
    s.run(6000, trace=False)
    s.reset_diagnostics()
    s.run(4000, 0, 40)                               # thinned trace of the window for the k histogram
    dg = s.diagnostics(allreduce=False)
    n = len(g["mean_k"])
    se = lambda a: a.std(ddof=1) / np.sqrt(n)
    # KCAP = LANES-1 = 15 changepoints per chain (the reference stores kmax = 10 but never
    # enforces it, changepoint.py:100): births proposed at k = 15 are rejected AND counted.
    assert dg["overflows"] < 1e-4 * K * 4000
    assert abs(dg["mean"][0] - g["mean_sig"].mean()) < 4 * se(g["mean_sig"]) + 1e-4      # sigma
    assert abs(dg["mean"][1] - g["mean_k"].mean()) < 4 * se(g["mean_k"])                 # E[k]
    assert abs(dg["accept_rate"] - g["accept"].mean()) < 4 * se(g["accept"]) + 2e-3
    hist_dev = np.bincount(s._chain_thetas.k.ravel(), minlength=16)[:16] / s._chain_thetas.k.size
    hist_ref, hist_se = g["khist"].mean(0), g["khist"].std(0, ddof=1) / np.sqrt(n)
    assert np.all(np.abs(hist_dev - hist_ref) < 4 * hist_se + 5e-3)


def test_full_size_self_consistency():
    """BASELINE config 2 size: 65,536 chains.  The log-posterior carried through T
    accept/reject steps must equal a fresh pointwise evaluation of the final states."""
    import torch
    from riemann_b200 import Sampler, _lib
    from riemann_b200.models.changepoint import ChangepointParams
    dm, dp, _, _ = _setup()
    K = 65536
    s = Sampler(dm, dp, ChangepointParams([2.0], [1.0, 3.0], 0.1), K=K, seed=7)
    s.run(300, trace=False)
    tr, lp = s._chain_thetas, np.asarray(s._chain_logpost[-1])
    assert np.all(np.isfinite(lp))
    dk, dx, dv, ds = (torch.as_tensor(a[-1], device="cuda") for a in (tr.k, tr.cpx, tr.cpv, tr.sig))
    out = torch.empty(K, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().rmn_model_cp_logpost(dm._handle, 0, K, _lib.ptr(dk), _lib.ptr(dx),
                                                _lib.ptr(dv), _lib.ptr(ds), _lib.ptr(out), _lib.stream_ptr()))
    assert relerr(out.cpu().numpy(), lp) < 1e-12
    k = tr.k[-1]
    assert k.min() >= 0 and k.max() < _lib.CP_LANES
    # sortedness / positivity invariants of every accepted state
    for c in np.random.default_rng(0).integers(0, K, 200):
        cx = tr.cpx[-1, c, :k[c]]
        assert np.all(np.diff(cx) > 0) and np.all(tr.cpv[-1, c, :k[c] + 1] > 0) and tr.sig[-1, c] > 0


def test_single_chain_warps_replay_the_reference_through_the_move_specific_guards(golden):
    """With K = 1 every group of the warp shadows the same chain, so the warp is uniform in the move type at every
    step and the kernel skips the blocks that move does not need (sigma move: no search, no sums; height move: cached
    gaps and run boundaries; location move: cached height prior).  The replay of the reference's stream must equal the
    fixture exactly as the 4-chains-in-one-warp replay does (which executes the union of all moves), bit for bit."""
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    g = golden("changepoint")
    dm, dp, _, _ = _setup(g)
    nch, T = g["tape"].shape[:2]
    th0 = [ChangepointParams(g["cpx"][c, 0, :g["k"][c, 0]], g["cpv"][c, 0, :g["k"][c, 0] + 1],
                             g["sig"][c, 0]) for c in range(nch)]
    s4 = Sampler(dm, dp, th0)
    ex4 = s4.run_injected(tape=np.transpose(g["tape"], (1, 0, 2)))
    for c in range(nch):
        s1 = Sampler(dm, dp, [th0[c]])
        ex1 = s1.run_injected(tape=g["tape"][c][:, None, :])
        tr = s1._chain_thetas
        assert np.array_equal(tr.k[:, 0], g["k"][c])
        assert relerr(tr.cpx[:, 0], g["cpx"][c]) < TOL and relerr(tr.cpv[:, 0], g["cpv"][c]) < TOL
        assert relerr(s1._chain_logpost[:, 0], g["logpost"][c]) < TOL
        assert relerr(ex1["prop_logpost"][:, 0], g["prop_logpost"][c]) < TOL
        assert np.array_equal(s1._chain_logpost[:, 0], s4._chain_logpost[:, c])          # same bits on both routes
        assert np.array_equal(ex1["prop_logpost"][:, 0], ex4["prop_logpost"][:, c], equal_nan=True)
        assert np.array_equal(ex1["accepted"][:, 0], ex4["accepted"][:, c])


@pytest.mark.parametrize("schedule,off", [("group", 0), ("group", 13), ("chain", 5)])
def test_time_sliced_launch_is_bit_identical(schedule, off, monkeypatch):
    """More blocks than the SMs hold at once and not a whole number of waves: the run goes through
    changepoint_sliced_kernel (persistent blocks, time slices of the launch striped over them, the chain state handed
    over through global memory).  States, log-posteriors and accept counts must equal those of the one-slice launch
    (RMN_CP_SLICED=0) bit for bit, across launches that continue each other; the diagnostics sums are added slice by
    slice, so they agree to rounding."""
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    dm, dp, _, _ = _setup()
    th0 = ChangepointParams([2.0], [1.0, 3.0], 0.1)
    K = 23000                                     # 719 blocks of 32 chains: more than 592 resident, not a multiple
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("RMN_CP_SLICED", mode)
        s = Sampler(dm, dp, th0, K=K, seed=3, chain_offset=off, move_schedule=schedule)
        s.run(450, trace=False)
        s.run(230, trace=False)                   # 2 slices of 115; continues from the first launch's state
        out[mode] = (s._download_state(), s.diagnostics(allreduce=False), s.chain_moments())
    (a_st, a_lp), a_d, a_m = out["0"]
    (b_st, b_lp), b_d, b_m = out["1"]
    for x, y in zip(a_st, b_st):
        assert np.array_equal(x, y)
    assert np.array_equal(a_lp, b_lp)
    assert a_d["accept_rate"] == b_d["accept_rate"] and a_d["overflows"] == b_d["overflows"]
    assert np.allclose(a_m[0], b_m[0], rtol=1e-12, atol=0) and np.allclose(a_m[1], b_m[1], rtol=1e-9, atol=1e-18)


def test_per_chain_schedule_without_guards_is_bit_identical(monkeypatch):
    """Per-chain move schedules run through an instantiation of the kernel with the warp-uniform guards compiled out
    (the warp executes the union of the moves anyway).  RMN_CP_GUARD=1 sends the same run through the guarded
    instantiation: states, log-posteriors and accept counts must agree bit for bit, one-slice and time-sliced."""
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    dm, dp, _, _ = _setup()
    th0 = ChangepointParams([2.0], [1.0, 3.0], 0.1)
    for K, T in ((203, 600), (23000, 240)):
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("RMN_CP_GUARD", mode)
            s = Sampler(dm, dp, th0, K=K, seed=9, chain_offset=3, move_schedule="chain")
            s.run(T, trace=False)
            out[mode] = (s._download_state(), s.diagnostics(allreduce=False))
        (a_st, a_lp), a_d = out["0"]
        (b_st, b_lp), b_d = out["1"]
        for x, y in zip(a_st, b_st):
            assert np.array_equal(x, y)
        assert np.array_equal(a_lp, b_lp)
        assert a_d["accept_rate"] == b_d["accept_rate"] and a_d["overflows"] == b_d["overflows"]


@pytest.mark.parametrize("schedule", ["group", "chain"])
def test_move_schedules_are_shard_invariant(schedule):
    """Philox mode: chains [off, off + n) of a sharded run equal the same chains of the unsharded run bit for bit, for
    both move schedules -- the schedule groups are aligned to GLOBAL chain ids, so a shard that starts in the middle
    of a group (off = 13) still shares its move draws with the right neighbours."""
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    dm, dp, _, _ = _setup()
    th0 = ChangepointParams([2.0], [1.0, 3.0], 0.1)
    full = Sampler(dm, dp, th0, K=96, seed=11, move_schedule=schedule)
    full.run(400, trace=False)
    (fk, fx, fv, fs), flp = full._download_state()
    for off, n in ((13, 30), (8, 16), (0, 5)):
        part = Sampler(dm, dp, th0, K=n, seed=11, chain_offset=off, move_schedule=schedule)
        part.run(400, trace=False)
        (pk, px, pv, ps), plp = part._download_state()
        assert np.array_equal(pk, fk[off:off + n]) and np.array_equal(px, fx[off:off + n])
        assert np.array_equal(pv, fv[off:off + n]) and np.array_equal(ps, fs[off:off + n])
        assert np.array_equal(plp, flp[off:off + n])


def test_shared_move_schedule_leaves_chain_means_uncorrelated():
    """rmn_sampler_set_move_schedule(1): the 8 chains of a group share the move-type draws.  Each chain is still an exact
    replica of the reference sampler; this checks the pooled estimators' premise -- the per-chain ergodic averages of
    a group are uncorrelated.  Intra-group correlation from Var(group mean) = Var(chain mean) (1 + 7 rho) / 8 over
    2,048 groups: |rho| < 4.5 standard errors for every tracked functional, and the pooled means of the two schedules
    agree."""
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams
    dm, dp, _, _ = _setup()
    K, G = 16384, 2048
    means = {}
    for schedule in ("group", "chain"):
        s = Sampler(dm, dp, ChangepointParams([2.0], [1.0, 3.0], 0.1), K=K, seed=5, move_schedule=schedule)
        s.run(30000, trace=False)
        s.reset_diagnostics()
        s.run(40000, trace=False)
        m, _ = s.chain_moments()
        se = np.sqrt(2.0 / (8 * 7 * G))
        for f in range(m.shape[0]):
            x = m[f]
            rho = (8.0 * x.reshape(G, 8).mean(axis=1).var(ddof=1) / x.var(ddof=1) - 1.0) / 7.0
            assert abs(rho) < 4.5 * se, (schedule, f, rho, se)
        means[schedule] = (m.mean(axis=1), m.std(axis=1, ddof=1) / np.sqrt(K))
    d = np.abs(means["group"][0] - means["chain"][0]) / np.hypot(means["group"][1], means["chain"][1])
    assert np.all(d < 4.5), d
