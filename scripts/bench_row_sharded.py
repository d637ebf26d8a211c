"""Strong scaling of the row-sharded data mode on BASELINE config 4's shape:
logistic regression N = 1e6, d = 100, MALA, K chains replicated on every rank, the N rows split
over the ranks, per-sweep partial sums all-reduced with NCCL inside rmn_sampler_run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/bench_row_sharded.py [--chains 1024] [--iters 2] [--steps 5]

Rank 0 first times the plain sampler on the full data set (one GPU), then all ranks time the
row-sharded sampler; CUDA events on the launching stream, max over ranks.  One JSON line."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(torch, s, iters, steps, warmup):
    for _ in range(warmup):
        s.run(iters, trace=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s.run(iters, trace=False)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=100)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    from riemann_b200 import Sampler
    from riemann_b200.distributed import shard_rows
    from riemann_b200.models.logistic import LogisticRegression
    from riemann_b200.proposals.hamiltonian import MALA

    from riemann_b200 import synthetic
    X, y, ts, pv = synthetic.logistic_problem(a.rows, a.dim)
    rng = np.random.Generator(np.random.Philox(synthetic.SEED_BASE + 5))
    th0 = ts[None, :] + 0.01 * rng.standard_normal((a.chains, a.dim))

    ms_plain = None
    if rank == 0:
        m = LogisticRegression(X, y, pv)
        s = Sampler(m, MALA(0.02, m.grad_log_posterior), th0, seed=1)
        ms_plain = timed(torch, s, a.iters, a.steps, a.warmup)
        ref_state = np.asarray(s.current_state()[0]).copy()
        del s, m
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    off, n = shard_rows(a.rows)
    m = LogisticRegression(X[off:off + n], y[off:off + n], pv)
    s = Sampler(m, MALA(0.02, m.grad_log_posterior), th0, seed=1, row_sharded=True)
    if world > 1:
        dist.barrier()
    ms = timed(torch, s, a.iters, a.steps, a.warmup)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        # same seed, same number of steps: the two samplers walked the same chains up to fp64 summation order
        drift = float(np.max(np.abs(np.asarray(s.current_state()[0]) - ref_state)))
        sweeps = a.chains * a.iters
        print(json.dumps({
            "metric": "mh_chain_steps_per_sec", "unit": "chain-steps/s", "n_gpus": world, "scaling": "strong",
            "config": {"workload": "logistic regression N=%d d=%d MALA, %d chains replicated, rows sharded over %d ranks"
                                   % (a.rows, a.dim, a.chains, world), "iters_per_step": a.iters},
            "value": sweeps / (ms * 1e-3), "ms_per_step": ms,
            "one_gpu_all_rows": {"value": sweeps / (ms_plain * 1e-3), "ms_per_step": ms_plain},
            "speedup_vs_one_gpu": ms_plain / ms, "steps": a.steps, "warmup": a.warmup,
            "max_abs_state_difference_vs_one_gpu": drift}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
