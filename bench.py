#!/usr/bin/env python
"""
bench.py -- MH chain-steps/sec (and min-ESS/sec) of the B200 engine on BASELINE.json's
metric, next to the CPU sampler timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Headline workload (N=1 default) is BASELINE.json configs[1]: the changepoint model of
examples/test_changepoint.py, 4-way RW/birth/death Metropolis proposal, 65,536 chains per
GPU.  One bench "step" = one launch of the hot path: T MH iterations of every chain
(`--iters`, default 1000), device-resident.  Chains shard over GPUs with no data-path
collective (weak scaling: 65,536 chains per GPU); the only exchange is the per-batch
diagnostics all-reduce (NCCL), which is inside the timed region.

Other workloads (`--workload gauss1000_mala | logistic_mala | logistic_mmala | gauss2d_rw`)
print the same JSON line for BASELINE configs 3, 4, 5 and 1, `gauss2d_pt` for the parallel-tempering
"next" row; `--precision tf32x3 | tf32-metric` selects the tcgen05 tensor-core modes of the dense
Gaussian and logistic workloads (default f64 = the parity mode).  They are for profiling and DESIGN.md;
the driver's line is the default one.

`roofline.achieved` is the algorithmic work of the timed region divided by the DOMINANT kernel's own time
in it (CUDA events around each of its launches, rmn_sampler_kernel_timing); `tau_check` (changepoint) puts
the Sokal-window autocorrelation time of a traced subset next to the moment-based one behind min_ess_per_sec.

Timing: W untimed warm-up steps, then K steps each bracketed by CUDA events on the
launching stream, L2 flushed (256 MiB write) between steps outside the event pairs,
barrier + synchronize on both sides, MAX over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mh_chain_steps_per_sec"
UNIT = "chain-steps/s"
CHAINS_PER_GPU = {"changepoint": 65536, "gauss2d_rw": 1 << 20, "gauss1000_mala": 16384,
                  "logistic_mala": 1024, "logistic_mmala": 512, "gauss2d_pt": 5 * (1 << 17)}
# SURVEY.md section 8d / BASELINE.md section 4: algorithmic work per chain-step
ALGO_FLOP = {"changepoint": 650.0, "gauss2d_rw": 40.0, "gauss1000_mala": 2.0e6,
             "logistic_mala": 4.0e8, "logistic_mmala": 4.4e8, "gauss2d_pt": 40.0}


# From the committed ncu captures (profiles/r1_*.md; `ncu --set full`, one launch of the dominant
# kernel on this code): DRAM bytes per launch and the pipe/issue utilisation.  Static evidence,
# NOT re-measured by this script -- the live numbers of a run are `value`, `ms_per_step`, `roofline.achieved`.
PROFILED = {
    ("gauss2d_rw", "f64"): {"kernel": "small_gauss_kernel<2,0,0,0>", "traffic": 92.7e6 + 36.6e6, "issue_slot_util": 0.652,
                            "fp64_pipe_active": 0.281, "source": "profiles/r1_gauss2d_small_gauss.md"},
    ("changepoint", "f64"): {"kernel": "changepoint_kernel<0,2,4>", "traffic": 27.58e6 + 0.07e6, "issue_slot_util": 0.673,
                             "fp64_pipe_active": 0.219, "warp_inst_per_chain_step": 159, "source": "profiles/r1_changepoint_gl4.md"},
    ("gauss1000_mala", "f64"): {"kernel": "gemm_abt_kernel<1>", "traffic": 455.6e6 + 125.1e6, "tensor_pipe_active": 0.770,
                                "source": "profiles/r1_gauss1000.md"},
    ("gauss1000_mala", "tf32x3"): {"kernel": "tf32x3_gemm_kernel<0,3>", "traffic": 144.8e6 + 42.0e6, "tensor_pipe_active": 0.855,
                                   "source": "profiles/r1_gauss1000_tf32x3_gemm.md"},
    ("logistic_mala", "f64"): {"kernel": "lg_eval_kernel", "traffic": 812.9e6 + 7.9e6, "tensor_pipe_active": 0.683,
                               "source": "profiles/r1_logistic.md"},
    ("logistic_mmala", "f64"): {"kernel": "lg_metric_kernel", "traffic": 149.1e6 + 104.5e6, "tensor_pipe_active": 0.687,
                                "source": "profiles/r1_mmala_metric_f64.md"},
    ("logistic_mmala", "tf32-metric"): {"kernel": "tf32x3_gemm_kernel<0,1> (metric GEMM; lg_eval_kernel is the larger share of the step)",
                                        "traffic": 3398.4e6 + 22.5e6, "tensor_pipe_active": 0.953,
                                        "source": "profiles/r1_mmala_metric_tf32_gemm.md"},
}


def load_peaks():
    """MEASURED_PEAKS.json (driver-written: HBM copy GB/s, bf16 cuBLAS TF/s) plus the figures it does
    not carry, measured here with the same method by scripts/peaks/measure_peaks.py and committed as
    profiles/measured_peaks_extra.json (fp64 DFMA / DMMA, TF32 cuBLAS).  Fallbacks are B200_PROFILING.md's."""
    d, src = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        src = "measured"
    x = os.path.join(ROOT, "profiles", "measured_peaks_extra.json")
    if os.path.exists(x):
        with open(x) as f:
            d["extra"] = json.load(f)
    return d, src


# ----------------------------------------------------------------------------
# CPU baseline: the numpy port of the reference Sampler (oracle/), one chain per core
# ----------------------------------------------------------------------------
def _cpu_worker(args):
    workload, seed, budget_s = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    from oracle import riemann_port as port
    np.random.seed(seed)
    if workload == "changepoint":
        model, prop, th0, _ = port.make_changepoint_problem()
        chunk = 500
    elif workload == "gauss2d_rw":
        model = port.benchmark_gauss(2)
        prop = port.MetropolisRandomWalk(0.5 * np.eye(2))
        th0 = np.ones(2)
        chunk = 2000
    elif workload == "gauss1000_mala":
        model = port.benchmark_gauss(1000)
        prop = port.MALA(0.08, model.grad_log_likelihood)
        th0 = np.zeros(1000)
        chunk = 5
    else:
        N, d = (100000, 64) if workload == "logistic_mmala" else (1000000, 100)
        X, y, ts, pv = port.make_logistic_problem(N, d)
        model = port.LogisticRegression(X, y, pv)
        prop = (port.SimplifiedMMALA(0.5, model) if workload == "logistic_mmala"
                else port.MALA(0.02, model.grad_log_posterior))
        th0 = ts.copy()
        chunk = 1
    s = port.Sampler(model, prop, th0)
    n = 0
    trace = []
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        while time.perf_counter() - t0 < budget_s:
            s.run(chunk)
            n += chunk
            if workload == "changepoint":
                trace.append(s._chain_thetas[-1].sig)
    return n, time.perf_counter() - t0


def cpu_baseline(workload, budget_s=12.0, cores=None):
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(workload, 1000 + i, budget_s) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return {"value": steps / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d independent chains (one per host core) of the numpy restatement of "
                      "riemann's Sampler.sample on the same synthetic %s problem, %.0f s each, "
                      "%d chain-steps total" % (cores, workload, budget_s, steps)}


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi sampled every 50 ms in the background; started BEFORE warm-up (the tool needs
    a few hundred ms to come up) and filtered to the wall-clock window of the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(gpu_index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]),
                             [nm for nm, v in zip(names, parts[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        self.f.close()
        os.unlink(self.f.name)
        inwin = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= self.t1 + 0.05]
        use = inwin
        if not use and rows and self.t0 is not None:          # window shorter than the sampling period
            mid = 0.5 * (self.t0 + self.t1)
            use = sorted(rows, key=lambda r: abs(r[0] - mid))[:2]
            out["note"] = "timed window shorter than the 50 ms sampling period: nearest samples used"
        if use:
            out.update(sm_mhz=float(np.median([r[1] for r in use])), sm_max_mhz=float(max(r[2] for r in use)),
                       samples=len(use), reasons=sorted({x for r in use for x in r[3]}))
        return out


# ----------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------
def build_workload(name, K, seed, chain_offset, precision="f64"):
    """Returns (sampler, host_inputs dict for the e2e leg, description)."""
    from riemann_b200 import Sampler, synthetic    # synthetic-input recipes; the engine arm never imports oracle/
    if name == "changepoint":
        from riemann_b200.models.changepoint import ChangepointParams, ChangepointRegression1D
        from riemann_b200.proposals.changepoint import ChangepointRegression1DProp
        c = synthetic.changepoint_problem()
        model = ChangepointRegression1D(c["x"], c["y"], c["xmin"], c["xmax"], c["lamb"], c["kmax"], c["alpha"], c["beta"])
        prop = ChangepointRegression1DProp(model, c["hscale"])
        s = Sampler(model, prop, ChangepointParams(*c["theta0"]), K=K, seed=seed,
                    chain_offset=chain_offset)
        return s, "changepoint regression (examples/test_changepoint.py recipe: M=100, 5 true changepoints)"
    if name == "gauss2d_rw":
        from riemann_b200.models import benchmarks
        from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
        s = Sampler(benchmarks.benchmark_gauss2d_corr, MetropolisRandomWalk(0.5 * np.eye(2)), np.ones(2),
                    K=K, seed=seed, chain_offset=chain_offset)
        return s, "benchmark_gauss2d_corr, MetropolisRandomWalk(0.5 I)"
    if name == "gauss2d_pt":
        # "next" row N3: parallel tempering, ladders of 5 temperatures (ptsampler.py default) along the chain axis
        from riemann_b200 import PTSampler
        from riemann_b200.models import benchmarks
        from riemann_b200.proposals.randomwalk import MetropolisRandomWalk
        if K % 5:
            raise SystemExit("gauss2d_pt: chains per GPU must be a multiple of 5 (one ladder = 5 temperatures)")
        pt = PTSampler(benchmarks.benchmark_gauss2d_corr, MetropolisRandomWalk(0.5 * np.eye(2)), np.ones(2),
                       K=K // 5, seed=seed, chain_offset=chain_offset // 5)
        pt._sampler._pt_owner = pt                      # keep the wrapper alive
        return pt._sampler, ("benchmark_gauss2d_corr, PTSampler: %d ladders x 5 temperatures (0.5**arange(5)), "
                             "Pswap = 0.1, MetropolisRandomWalk(0.5 I)" % (K // 5))
    if name == "gauss1000_mala":
        from riemann_b200.models import benchmarks
        from riemann_b200.proposals.hamiltonian import MALA
        m = benchmarks.gauss_corr(1000)
        rng = np.random.Generator(np.random.Philox(synthetic.SEED_BASE + 3))
        th0 = rng.standard_normal((K, 1000))
        s = Sampler(m, MALA(0.08, m.grad_log_likelihood), th0, seed=seed, chain_offset=chain_offset,
                    precision=precision)
        return s, "dense Gaussian d=1000 (0.1 I + 0.9 11^T), MALA eps=0.08, precision " + precision
    if name in ("logistic_mala", "logistic_mmala"):
        from riemann_b200.models.logistic import LogisticRegression
        from riemann_b200.proposals.hamiltonian import MALA, SimplifiedMMALA
        N, d = (100000, 64) if name == "logistic_mmala" else (1000000, 100)
        X, y, ts, pv = synthetic.logistic_problem(N, d)
        m = LogisticRegression(X, y, pv)
        rng = np.random.Generator(np.random.Philox(synthetic.SEED_BASE + 5))
        th0 = ts[None, :] + 0.01 * rng.standard_normal((K, d))
        prop = SimplifiedMMALA(0.5, m) if name == "logistic_mmala" else MALA(0.02, m.grad_log_posterior)
        if precision == "tf32-metric" and name != "logistic_mmala":
            raise SystemExit("--precision tf32-metric applies to logistic_mmala only")
        s = Sampler(m, prop, th0, seed=seed, chain_offset=chain_offset, precision=precision)
        return s, "logistic regression N=%d d=%d, %s%s" % (
            N, d, "simplified mMALA" if "mm" in name else "MALA",
            {"f64": "", "tf32-metric": ", Fisher metric on tcgen05 (TF32 GEMM)",
             "tf32x3": ", likelihood sweep on tcgen05 (3xTF32 GEMMs, fp64 pointwise stage)"}[precision])
    raise SystemExit("unknown workload %r" % name)


def run_engine(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0 and not os.path.exists(ge.OUT):
        ge.build()
    if world > 1:
        dist.barrier()
    from riemann_b200 import _lib
    from riemann_b200.distributed import reduce_block, summarize_block, summarize_split

    wl = args.workload
    Kg = args.chains or CHAINS_PER_GPU[wl]
    T = args.iters or {"changepoint": 1000, "gauss2d_rw": 2000, "gauss1000_mala": 20,
                       "logistic_mala": 2, "logistic_mmala": 2, "gauss2d_pt": 2000}[wl]
    seed = 20261018
    s, desc = build_workload(wl, Kg, seed, rank * Kg, args.precision)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # burn-in (untimed setup), then warm-up steps
    clocks = ClockSampler(local) if rank == 0 else None
    burn = args.burn if args.burn is not None else {"changepoint": 10000}.get(wl, 2 * T)
    s.run(burn, trace=False)
    for _ in range(max(args.warmup, 3)):
        s.run(T, trace=False)
        blk = reduce_block(s.diagnostics_block())
    s.reset_diagnostics()
    sync_all()

    lib = _lib.load()
    launches0 = s.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    s.enable_kernel_timing(True)                   # CUDA events around the dominant kernel's launches
    sync_all()
    if clocks:
        clocks.begin()
    half = args.steps // 2 if args.steps >= 2 and args.steps % 2 == 0 else 0
    blk_first = None
    for i in range(args.steps):
        flush.fill_(i & 0xff)                      # L2 flush, outside the event pair
        if half and i == half:                     # second half-window of the split-chain diagnostics
            blk_first = blk.clone()
            s.reset_diagnostics()
        ev[i][0].record(stream)
        _lib.check(lib.rmn_sampler_run(s._handle, T, None, None, _lib.stream_ptr()))
        s.total_steps += T
        blk = reduce_block(s.diagnostics_block())  # per-batch diagnostics all-reduce (NCCL)
        ev[i][1].record(stream)
    sync_all()
    if clocks:
        clocks.end()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    kt = s.kernel_timing()                         # dominant kernel only: total ms / launches in the timed region
    s.enable_kernel_timing(False)
    clk = clocks.stop() if clocks else None
    launches = s.launch_count - launches0
    t_all = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms = float(t_all.item())
    diag = None
    if blk_first is not None:
        # split-chain diagnostics over the two halves of the timed region (split-R-hat, ESS of 2K half-chains)
        try:
            diag = summarize_split(blk_first.cpu().numpy(), blk.cpu().numpy())
            diag_kind = "split: %d half-chains of %d MH steps" % (diag["chains"], diag["steps"])
        except ValueError:
            diag = None
    if diag is None:
        diag = summarize_block(blk.cpu().numpy())
        diag_kind = "whole window"
    K_total = Kg * world
    value = K_total * T * args.steps / (ms * 1e-3)

    # ---- end-to-end leg: host buffers in, host results out, copies inside the timed region
    e2e = run_e2e(args, s, T, Kg, world, sync_all)

    # ---- estimator cross-check (outside every timed region): Sokal-window integrated autocorrelation time
    #      (emcee's estimator, examples/test_randomwalk.py:42) of a traced subset of chains next to the
    #      moment-based tau = n B / W of the diagnostics block that min_ess_per_sec is computed from
    tau_check = None
    if rank == 0 and wl == "changepoint":
        from riemann_b200 import diagnostics as dgn
        s2, _ = build_workload(wl, 128, seed + 1, 0, args.precision)
        s2.run(burn, trace=False)
        s2.run(8000, 0, 1)
        tr = s2._chain_thetas
        x = np.stack([tr.sig[1:], tr.k[1:].astype(np.float64)], axis=2)          # [N, K, 2]: sigma, k
        # RMN_BENCH_DEVICE_TAU=1: the same estimator on the GPU (rmn_autocorr_tau); host FFT by default
        tau_s = dgn.integrated_time_chains(x, device=os.environ.get("RMN_BENCH_DEVICE_TAU", "0") == "1")
        tau_m = diag["tau"][:2]
        tau_check = {"functionals": ["sigma", "k"], "sokal_tau_steps": [float(t) for t in tau_s],
                     "moment_tau_steps": [float(t) for t in tau_m],
                     "note": "Sokal tau (c = 5) of 128 traced chains x 8000 steps vs tau = n B / W of the "
                             "%d bench chains (MH steps per independent sample)" % K_total}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    ms_launch = ms / args.steps
    # roofline.achieved = ALGORITHMIC work of the timed region / the dominant kernel's OWN time in it
    # (CUDA events on the launching stream, rmn_sampler_kernel_timing); the kernel's share of the step is
    # reported next to it and must agree with the ncu launch list under profiles/.
    algo_flop_step = ALGO_FLOP[wl] * Kg * T              # per bench step, per GPU
    if wl == "logistic_mmala" and args.precision == "tf32-metric":
        algo_flop_step = 4.0 * 100000 * 64 * Kg * T      # lg_eval_kernel's part: logits + gradient, 4 N d
    kernel_ms = kt["total_ms"] if kt["launches"] > 0 else ms
    ach_tflops = algo_flop_step * args.steps / (kernel_ms * 1e-3) / 1e12
    extra = peaks.get("extra", {})
    computed_fp64 = 148 * 64 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12      # fp64 FMA lanes x clock
    if wl in ("changepoint", "gauss2d_rw", "gauss2d_pt"):
        peak = extra.get("fp64_dfma_tflops", computed_fp64)
        roof = {"bound": "alu", "achieved": ach_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": ach_tflops / peak, "traffic": None,
                "note": "issue-bound fp64/integer kernel (SURVEY 8d: not HBM, not tensor); achieved = "
                        "%.0f algorithmic op/chain-step x rate; peak = fp64 DFMA rate %s; the binding resource "
                        "is the warp-instruction issue rate, see `profiled` (issue-slot utilisation from ncu)"
                        % (ALGO_FLOP[wl], "measured by scripts/peaks (profiles/measured_peaks_extra.json)"
                           if "fp64_dfma_tflops" in extra else "computed as 148 SM x 64 lanes x 2 x clocks.max.sm")}
    elif args.precision == "tf32-metric" and wl == "logistic_mmala":
        peak = extra.get("fp64_dmma_tflops", computed_fp64)
        roof = {"bound": "tensor", "achieved": ach_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": ach_tflops / peak, "traffic": None,
                "note": "dominant kernel = lg_eval_kernel (fp64 DMMA, 4 N d = 2.6e7 flop per chain-step) against the "
                        "measured DMMA peak; the metric GEMM (4.2e8 flop per chain-step) runs on tcgen05 in TF32 at 95 % "
                        "tensor-pipe utilisation (profiles/r1_mmala_metric_tf32_gemm.md) and is the smaller share of the step"}
    elif args.precision == "tf32x3" and wl.startswith("logistic"):
        peak = extra.get("tf32_cublas_tflops", 0.5 * peaks["bf16_tflops"])
        n_, d_ = (100000, 64) if wl == "logistic_mmala" else (1000000, 100)
        ach_tflops = 4.0 * n_ * d_ * Kg * T * args.steps / (kernel_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": ach_tflops / peak, "traffic": None,
                "note": "likelihood sweep = logits GEMM + fp64 pointwise kernel + split-K gradient GEMM, timed together "
                        "(CUDA events around the three launches); achieved counts the ALGORITHMIC 4 N d flop per chain-step "
                        "once (3 TF32 MMAs are issued per product, and the d-wide gradient tile fills 100/256 resp. 64/256 "
                        "of the MMA's N), so the fraction of the cuBLAS TF32 peak is small by construction; the sweep is "
                        "bounded by the fp64 pointwise stage and the 12 B/(chain,row) of Z/R traffic"}
    elif args.precision == "tf32x3":
        peak = extra.get("tf32_cublas_tflops", 0.5 * peaks["bf16_tflops"])
        roof = {"bound": "tensor", "achieved": ach_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": ach_tflops / peak, "traffic": None,
                "note": "tcgen05 kind::tf32, 3 MMAs per product (fp32-accurate): achieved counts the ALGORITHMIC "
                        "2 d^2 flop per chain-step once, so 1/3 is the ceiling of this scheme; peak = %s"
                        % ("cuBLAS TF32 8192^3 measured by scripts/peaks (profiles/measured_peaks_extra.json)"
                           if "tf32_cublas_tflops" in extra else "half the %s bf16 figure of MEASURED_PEAKS.json" % peak_src)}
    else:
        peak = extra.get("fp64_dmma_tflops", computed_fp64)
        roof = {"bound": "tensor", "achieved": ach_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": ach_tflops / peak, "traffic": None,
                "note": "fp64 tensor path (DMMA m8n8k4); peak = %s (bf16 %s peak %.0f TF/s for context only)"
                        % ("mma.sync.m8n8k4.f64 rate measured by scripts/peaks (profiles/measured_peaks_extra.json)"
                           if "fp64_dmma_tflops" in extra else "computed 148 SM x 64 lanes x 2 x clocks.max.sm",
                           peak_src, peaks["bf16_tflops"])}
    roof["kernel"] = kt["kernel"]
    roof["kernel_launches"] = kt["launches"]
    roof["kernel_ms_per_launch"] = kernel_ms / max(kt["launches"], 1)
    roof["kernel_share_of_step"] = kernel_ms / ms if ms > 0 else None
    prof = PROFILED.get((wl, args.precision))
    if prof:
        roof["traffic"] = prof["traffic"]
        roof["profiled"] = prof
    cpu = None if args.no_cpu else cpu_baseline(wl, args.cpu_seconds)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_launch, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"f64": "f64", "tf32x3": ("tf32x3 likelihood GEMMs, f64 pointwise + accept test" if wl.startswith("logistic")
                                                                       else "tf32x3/f32 state, f64 accept test"),
                                                            "tf32-metric": "f64 (proposal metric: tf32 tensor cores)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "%s: %s; %d chains/GPU x %d MH iterations per step; Philox4x32-10 RNG; "
                               "L2 flushed (256 MiB write) between steps, per-step CUDA events summed"
                               % (wl, desc, Kg, T),
                   "chains_per_gpu": Kg, "chains_total": K_total, "iters_per_step": T,
                   "burn_in_iters": burn, "parallelism": "chains sharded, dp%d" % world},
        "min_ess_per_sec": (diag["min_ess"] * 1.0) / (ms * 1e-3) if np.isfinite(diag["min_ess"]) else None,
        "diagnostics": {"estimator": diag_kind, "accept_rate": diag["accept_rate"], "max_rhat": float(np.nanmax(diag["rhat"])),
                        "min_ess": diag["min_ess"], "overflows": diag["overflows"],
                        "chains": diag["chains"], "steps_per_chain": diag["steps"]},
        "tau_check": tau_check,
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        "clocks": clk, "peaks_source": peak_src,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, s, T, Kg, world, sync_all):
    """Same metric through the public API with HOST buffers: every step uploads the chains'
    start states from pinned host memory, runs T iterations, and reads states, log-posteriors
    and the diagnostics block back to the host."""
    import torch
    import torch.distributed as dist
    from riemann_b200 import _lib
    lib = _lib.load()
    steps = max(2, min(args.steps, 5))
    if s._is_cp:
        (k, cpx, cpv, sig), lp = s._download_state()
        host_in = [torch.from_numpy(a).pin_memory() for a in (k, cpx, cpv, sig)]
        dev_in = [torch.empty_like(h, device="cuda") for h in host_in]
        dev_out = [torch.empty_like(h, device="cuda") for h in host_in] + \
                  [torch.empty(Kg, dtype=torch.float64, device="cuda")]
    else:
        th, lp = s._download_state()
        host_in = [torch.from_numpy(th).pin_memory()]
        dev_in = [torch.empty_like(host_in[0], device="cuda")]
        dev_out = [torch.empty_like(host_in[0], device="cuda"), torch.empty(Kg, dtype=torch.float64, device="cuda")]
    host_out = [torch.empty_like(d, device="cpu").pin_memory() for d in dev_out]
    nd = lib.rmn_sampler_diag_dim(s._handle)
    host_blk = torch.empty(_lib.DIAG_HDR + 3 * nd, dtype=torch.float64).pin_memory()
    h2d = sum(h.numel() * h.element_size() for h in host_in)
    d2h = sum(h.numel() * h.element_size() for h in host_out) + host_blk.numel() * 8

    def one():
        for h, d in zip(host_in, dev_in):
            d.copy_(h, non_blocking=True)
        if s._is_cp:
            _lib.check(lib.rmn_sampler_cp_set_state(s._handle, *[_lib.ptr(d) for d in dev_in], _lib.stream_ptr()))
        else:
            _lib.check(lib.rmn_sampler_set_state(s._handle, _lib.ptr(dev_in[0]), _lib.stream_ptr()))
        _lib.check(lib.rmn_sampler_run(s._handle, T, None, None, _lib.stream_ptr()))
        if s._is_cp:
            _lib.check(lib.rmn_sampler_cp_get_state(s._handle, *[_lib.ptr(d) for d in dev_out], _lib.stream_ptr()))
        else:
            _lib.check(lib.rmn_sampler_get_state(s._handle, _lib.ptr(dev_out[0]), _lib.ptr(dev_out[1]), _lib.stream_ptr()))
        blk = s.diagnostics_block()
        for h, d in zip(host_out, dev_out):
            h.copy_(d, non_blocking=True)
        host_blk.copy_(blk, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    one()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    sync_all()
    dt = time.perf_counter() - t0
    t_all = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    dt = float(t_all.item())
    return {"value": Kg * world * T * steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps,
            "note": "host wall clock around upload(pinned) -> set_state -> run -> get_state -> download"}


def run_reference(args):
    """--impl reference: the CPU sampler on this box's host cores, same metric/config.
    The reference is pure Python and cannot travel to the GPU box, so this arm times the
    numpy port (oracle/riemann_port.py, pinned to the reference's chains by tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    W, K = max(args.warmup, 1), max(args.steps, 1)
    per = max(2.0, min(20.0, 60.0 / (W + K)))
    for _ in range(1):
        cpu_baseline(wl, 1.0)                       # warm the pool / imports
    vals = [cpu_baseline(wl, per) for _ in range(K)]
    v = float(np.mean([c["value"] for c in vals]))
    cb = dict(vals[-1], value=v)
    Kg = args.chains or CHAINS_PER_GPU[wl]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (CPU reference arm: one chain per host core, %.0f s per step)" % (wl, per),
                       "chains_per_gpu": Kg},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="changepoint", choices=sorted(CHAINS_PER_GPU))
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU")
    ap.add_argument("--iters", type=int, default=None, help="MH iterations per step")
    ap.add_argument("--burn", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--precision", default="f64", choices=["f64", "tf32x3", "tf32-metric"],
                    help="tf32x3: tcgen05 tensor-core mode of the dense Gaussian workload; tf32-metric: "
                         "tcgen05 Fisher-metric GEMM of the logistic mMALA workload")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
