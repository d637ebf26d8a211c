"""
Row-sharded data mode of the logistic family (include/riemann_b200.h rmn_sampler_set_row_comm):
every rank holds a slice of the data rows and the same K chains; the per-chain partial log-likelihoods, gradients
and metrics are all-reduced with NCCL inside rmn_sampler_run.

  * one GPU: a world-size-1 communicator takes the full exchange path (fold of the row splits, grouped all-reduce,
    finish kernels on the folded sums) and must reproduce the plain sampler BIT FOR BIT;
  * two GPUs (skipped on a one-GPU box): the rows of the fixtures' data set are split over two ranks and the
    injected reference stream is replayed; both ranks must agree bit for bit with each other and to 1e-9 / 1e-8
    (fp64 sums in a different order) with the chain recorded through the reference's Sampler.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from gpu_helpers import relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _proposal(name, g, dm):
    from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA, VanillaHMC
    if name == "mala_logistic":
        return MALA(float(g["eps"]), dm.grad_log_posterior)
    if name == "hmc3_logistic":
        return VanillaHMC(float(g["eps"]), int(g["nsteps"]), dm.grad_log_posterior)
    return SimplifiedMMALA(float(g["eps"]), dm)


@pytest.mark.parametrize("name", ["mala_logistic", "hmc3_logistic", "mmala_logistic"])
def test_world1_communicator_is_bitwise_the_plain_sampler(golden, name):
    from riemann_b200 import Sampler
    from riemann_b200.models.logistic import LogisticRegression
    g = golden(name)
    dm = LogisticRegression(g["X"], g["y"], float(g["prior_var"]))
    out = []
    for sharded in (False, True):
        s = Sampler(dm, _proposal(name, g, dm), g["thetas"][0], row_sharded=sharded)
        s.run_injected(xi=g["xi"], u=g["u"])
        out.append((np.array(s._chain_thetas), np.array(s._chain_logpost)))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert relerr(out[1][0], g["thetas"]) < 1e-8


def test_philox_many_chains_world1_bitwise():
    """K = 70 chains, several row splits, Philox mode: the exchange path changes nothing at world size 1."""
    from oracle import riemann_port as port
    from riemann_b200 import Sampler
    from riemann_b200.models.logistic import LogisticRegression
    from riemann_b200.proposals.hamiltonian import AdaptScaleHMC
    X, y, ts, pv = port.make_logistic_problem(3000, 10, seed=5)
    dm = LogisticRegression(X, y, pv)
    th0 = ts[None] + 0.1 * np.random.default_rng(0).standard_normal((70, 10))
    res = []
    for sharded in (False, True):
        s = Sampler(dm, AdaptScaleHMC(0.05, 2, dm.grad_log_posterior), th0, seed=9, row_sharded=sharded)
        s.run(30, trace=False)
        res.append(np.asarray(s.current_state()[0]))
    assert np.array_equal(res[0], res[1])


def test_row_sharding_is_refused_where_it_does_not_exist(golden):
    from riemann_b200 import ParameterError, Sampler
    from riemann_b200.models.benchmarks import benchmark_gauss2d_corr
    from riemann_b200.models.logistic import LogisticRegression
    from riemann_b200.proposals.hamiltonian import MALA
    from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
    with pytest.raises((ParameterError, RuntimeError)):
        Sampler(benchmark_gauss2d_corr, MetropolisRandomWalk(np.eye(2)), np.ones(2), row_sharded=True)
    g = golden("mala_logistic")
    dm = LogisticRegression(g["X"], g["y"], float(g["prior_var"]))
    with pytest.raises((ParameterError, RuntimeError)):
        Sampler(dm, MALA(float(g["eps"]), dm.grad_log_posterior), g["thetas"][0], precision="tf32x3", row_sharded=True)


_TWO_RANK = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
from riemann_b200 import Sampler
from riemann_b200.distributed import shard_rows
from riemann_b200.models.logistic import LogisticRegression
from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA, VanillaHMC
from gpu_helpers import relerr
for name, tol in (("mala_logistic", 1e-9), ("hmc3_logistic", 1e-9), ("mmala_logistic", 1e-8)):
    g = np.load(os.path.join(sys.argv[1], "tests", "golden", name + ".npz"))
    off, n = shard_rows(len(g["y"]))
    dm = LogisticRegression(g["X"][off:off + n], g["y"][off:off + n], float(g["prior_var"]))
    if name == "mala_logistic": p = MALA(float(g["eps"]), dm.grad_log_posterior)
    elif name == "hmc3_logistic": p = VanillaHMC(float(g["eps"]), int(g["nsteps"]), dm.grad_log_posterior)
    else: p = SimplifiedMMALA(float(g["eps"]), dm)
    s = Sampler(dm, p, g["thetas"][0], row_sharded=True)
    ex = s.run_injected(xi=g["xi"], u=g["u"])
    th = np.array(s._chain_thetas)
    assert relerr(th, g["thetas"]) < tol, (name, relerr(th, g["thetas"]))
    assert relerr(s._chain_logpost, g["logpost"]) < tol
    assert np.array_equal(ex["accepted"][:, 0], np.any(g["thetas"][1:] != g["thetas"][:-1], axis=1))
    mine = torch.as_tensor(th, device="cuda"); other = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    assert all(torch.equal(o, mine) for o in other), name + ": ranks disagree"
    if rank == 0: print("row-sharded %s over %d ranks: rel err %.2e, ranks bit-identical" % (name, world, relerr(th, g["thetas"])), flush=True)
dist.destroy_process_group()
'''


def test_two_ranks_split_the_rows(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script), ROOT],
                       capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ranks bit-identical") == 3, r.stdout
