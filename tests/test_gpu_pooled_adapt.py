"""
GPU: covariance adaptation pooled over the chains (riemann_b200.proposals.randomwalk.PooledAdaptCovRandomWalk,
rmn_proposal_rw_set_pooled_cov_adapt).  The reference adapts from one chain's history (adaptive.py:38-103); there is no
reference stream to replay for the pooled form, so the checks are the ones the method itself promises: the pooled estimate
converges to the target covariance, the adapted factor gives the acceptance rate random-walk theory predicts for
C = 2.38^2 / d * Sigma (about 0.23-0.35 in moderate d), the chains keep sampling the target, and the device's estimate
equals the numpy estimate of the same pooled states.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _target(d, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    S = A @ A.T / d + 0.2 * np.eye(d)
    s = np.exp(rng.uniform(-1.5, 1.5, d))              # scales spread over a factor of 20
    return S * np.outer(s, s), rng.standard_normal(d)


@pytest.mark.parametrize("d,K", [(12, 2048), (37, 4096)])
def test_pooled_estimate_converges_and_tunes_the_walk(d, K):
    from riemann_b200 import Sampler
    from riemann_b200.models.gaussian import MultiGaussianDist
    from riemann_b200.proposals.randomwalk import PooledAdaptCovRandomWalk, MetropolisRandomWalk
    Sigma, mu = _target(d, d)
    m = MultiGaussianDist(mu, Sigma)
    rng = np.random.default_rng(1)
    th0 = mu + 0.1 * rng.standard_normal((K, d))                         # start far too narrow
    p = PooledAdaptCovRandomWalk(1e-4 * np.eye(d), t_adapt=25, stop_after=3000)
    s = Sampler(m, p, th0, seed=3)
    s.run(3000, trace=False)
    # adaptations happen before steps 25, 50, ..., 3000 are proposed; step 3000 belongs to the next call
    assert p.pool_count == K * (3000 // 25 - 1) and p.C.shape == (d, d)
    s.reset_diagnostics()
    s.run(1500, trace=False)                                             # one last adaptation, then plain MH
    assert p.pool_count == K * (3000 // 25)                              # ... and none after stop_after
    dg = s.diagnostics(allreduce=False)
    assert 0.15 < dg["accept_rate"] < 0.45, dg["accept_rate"]
    # the tuned walk samples the target: moments of the final states over the K chains
    th = np.asarray(s._chain_thetas[-1])
    assert np.max(np.abs(th.mean(axis=0) - mu) / np.sqrt(np.diag(Sigma))) < 6.0 / np.sqrt(K)
    emp = np.cov(th.T)
    assert np.max(np.abs(emp - Sigma) / np.sqrt(np.outer(np.diag(Sigma), np.diag(Sigma)))) < 0.15
    # the pooled estimate is dominated by the (many) post-transient samples: close to the target as well
    rel = np.abs(p.C - Sigma) / np.sqrt(np.outer(np.diag(Sigma), np.diag(Sigma)))
    assert np.max(rel) < 0.25, np.max(rel)
    # an untuned walk of the same starting covariance barely moves in the same number of steps
    s0 = Sampler(m, MetropolisRandomWalk(1e-4 * np.eye(d)), th0, seed=3)
    s0.run(3000, trace=False)
    th_untuned = np.asarray(s0._chain_thetas[-1])
    assert np.mean(np.var(th_untuned, axis=0) / np.diag(Sigma)) < 0.5 * np.mean(np.var(th, axis=0) / np.diag(Sigma))


def test_device_estimate_equals_numpy_on_the_same_states():
    """One adaptation from a known population: t_adapt = 1 adapts before step 1 is proposed, i.e. on the states after
    step 0 -- which are the traced record 1."""
    from riemann_b200 import Sampler
    from riemann_b200.models.gaussian import MultiGaussianDist
    from riemann_b200.proposals.randomwalk import PooledAdaptCovRandomWalk
    d, K = 20, 777
    Sigma, mu = _target(d, 5)
    m = MultiGaussianDist(mu, Sigma)
    th0 = mu + np.random.default_rng(2).standard_normal((K, d))
    p = PooledAdaptCovRandomWalk(0.01 * np.eye(d), t_adapt=1, stop_after=1)
    s = Sampler(m, p, th0, seed=9)
    s.run(2)
    assert p.pool_count == K
    pooled = np.asarray(s._chain_thetas[1])
    ref_mean = pooled.mean(axis=0)
    ref_cov = (pooled - ref_mean).T @ (pooled - ref_mean) / K
    assert np.max(np.abs(p.pool_mean - ref_mean)) < 1e-10
    assert np.max(np.abs(p.C - ref_cov)) < 1e-9 * np.max(np.abs(ref_cov))


def test_refusals():
    from riemann_b200 import Sampler
    from riemann_b200.models.gaussian import MultiGaussianDist
    from riemann_b200.proposals.randomwalk import PooledAdaptCovRandomWalk
    from riemann_b200.sampling_errors import ParameterError
    m2 = MultiGaussianDist(np.zeros(2), np.eye(2))
    with pytest.raises(ParameterError):
        Sampler(m2, PooledAdaptCovRandomWalk(np.eye(2)), np.zeros(2))               # small-d path: per-chain AdaptCov
    with pytest.raises(ParameterError):
        PooledAdaptCovRandomWalk(np.eye(12), t_adapt=0)
    m12 = MultiGaussianDist(np.zeros(12), np.eye(12))
    s = Sampler(m12, PooledAdaptCovRandomWalk(np.eye(12)), np.zeros((4, 12)))
    with pytest.raises(ParameterError):
        s.get_checkpoint()


_TWO_RANK = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
from riemann_b200 import Sampler
from riemann_b200.models.gaussian import MultiGaussianDist
from riemann_b200.proposals.randomwalk import PooledAdaptCovRandomWalk
d, K = 16, 1024
rng = np.random.default_rng(7)
A = rng.standard_normal((d, d)); Sigma = A @ A.T / d + 0.3 * np.eye(d); mu = rng.standard_normal(d)
m = MultiGaussianDist(mu, Sigma)
th0 = mu + 0.1 * np.random.default_rng(100 + rank).standard_normal((K, d))
p = PooledAdaptCovRandomWalk(1e-3 * np.eye(d), t_adapt=20, stop_after=2000)
s = Sampler(m, p, th0, seed=3, chain_offset=rank * K)
s.run(2000, trace=False); s.run(20, trace=False)
assert p.pool_count == world * K * 100, p.pool_count
mine = torch.as_tensor(p.C, device="cuda"); other = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(other, mine)
assert all(torch.equal(o, mine) for o in other), "ranks hold different pooled covariances"
rel = np.max(np.abs(p.C - Sigma) / np.sqrt(np.outer(np.diag(Sigma), np.diag(Sigma))))
assert rel < 0.25, rel
if rank == 0: print("pooled covariance over %d ranks: %d samples, max rel err %.3f, ranks bit-identical" % (world, p.pool_count, rel), flush=True)
dist.destroy_process_group()
'''


def test_two_ranks_share_one_pool(tmp_path):
    """The ranks' chains join ONE pool (NCCL all-reduce of the running sums inside rmn_sampler_run): every rank ends up
    with the same covariance estimate, built from world x K samples per adaptation."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "two_rank_pool.py"
    script.write_text(_TWO_RANK)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29579", str(script), root],
                       capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ranks bit-identical" in r.stdout, r.stdout
