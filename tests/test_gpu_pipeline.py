"""
GPU: riemann_b200.pipeline.HostJobRunner -- independent jobs from pinned host buffers through one device sampler.
The pipelined order of copies and kernels must not change any result: the overlapped runner returns, job by job, exactly
what the plain upload -> run -> download sequence returns (same sampler seed, same job order).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gauss_sampler(K, d, seed):
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    m = benchmarks.gauss_corr(d)
    return Sampler(m, MetropolisRandomWalk(0.05 * np.eye(d)), np.zeros((K, d)), seed=seed)


@pytest.mark.parametrize("d", [2, 40])
def test_overlapped_jobs_equal_serial_jobs(d):
    import torch
    from riemann_b200.pipeline import HostJobRunner
    K, T, njobs = 512, 50, 5
    rng = np.random.default_rng(3)
    jobs = [(torch.from_numpy(rng.standard_normal((K, d))).pin_memory(),) for _ in range(njobs)]
    got = {}
    for overlap in (False, True):
        s = _gauss_sampler(K, d, seed=11)
        r = HostJobRunner(s, overlap=overlap)
        assert r.host_layout() == [((K, d), torch.float64)]
        res = []
        for out in r.run(jobs, T):
            res.append((out["state"][0].numpy().copy(), out["logpost"].numpy().copy(), out["diagnostics"].numpy().copy()))
        got[overlap] = res
        assert len(res) == njobs
    for a, b in zip(got[False], got[True]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # different jobs do give different chains (the comparison above is not vacuous), and every job moved
    assert not np.array_equal(got[True][0][0], got[True][1][0])
    assert not np.array_equal(got[True][0][0], jobs[0][0].numpy())


def test_changepoint_jobs(golden):
    """The variable-dimension model: jobs are (k, cpx, cpv, sig) tuples; a job started from the fixture's states gives
    the same final states through the runner as through Sampler.run."""
    import torch
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.models.changepoint import ChangepointParams, ChangepointRegression1D
    from riemann_b200.proposals.changepoint import ChangepointRegression1DProp
    from riemann_b200.pipeline import HostJobRunner
    pm, pprop, pth0, _ = port.make_changepoint_problem()
    model = ChangepointRegression1D(pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb, pm.kmax, pm.alpha, pm.beta)
    th0 = ChangepointParams(pth0.cpx, pth0.cpv, pth0.sig)
    K, T = 256, 200
    ref = Sampler(model, ChangepointRegression1DProp(model, pprop.hscale), th0, K=K, seed=5)
    (k, cpx, cpv, sig), _ = ref._download_state()
    ref.run(T, trace=False)
    (k1, cpx1, cpv1, sig1), lp1 = ref._download_state()
    s = Sampler(model, ChangepointRegression1DProp(model, pprop.hscale), th0, K=K, seed=5)
    job = tuple(torch.from_numpy(a).pin_memory() for a in (k, cpx, cpv, sig))
    outs = list(HostJobRunner(s).run([job], T))
    assert np.array_equal(outs[0]["state"][0].numpy(), k1)
    assert np.array_equal(outs[0]["state"][3].numpy(), sig1)
    assert np.array_equal(outs[0]["logpost"].numpy(), lp1)
