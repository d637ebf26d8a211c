import numpy as np, sys, os, torch
sys.path.insert(0, "/root/repo")
from oracle import riemann_port as port
from riemann_b200 import Sampler, _lib
from riemann_b200.models.changepoint import ChangepointParams, ChangepointRegression1D, unpack_state
from riemann_b200.proposals.changepoint import ChangepointRegression1DProp
pm, pp, th0, _ = port.make_changepoint_problem()
m = ChangepointRegression1D(pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb, pm.kmax, pm.alpha, pm.beta)
p = ChangepointRegression1DProp(m, pp.hscale)
z = np.load("/root/repo/scratch/tpc_state.npz")
K, step, seed = 64, int(z["step"]), 2024
lib = _lib.load()
# device draws of that step
nz = torch.empty((K, 16), dtype=torch.float64, device="cuda"); uu = torch.empty(K, dtype=torch.float64, device="cuda")
_lib.check(lib.rmn_rng_draws(seed, 0, step, K, 16, _lib.ptr(nz), _lib.ptr(uu), _lib.stream_ptr()))
def raw(blk):
    ctr = np.zeros((K, 4), dtype=np.uint32); ctr[:, 0] = blk; ctr[:, 1] = step; ctr[:, 3] = np.arange(K)
    key = np.tile(np.array([seed & 0xffffffff, seed >> 32], dtype=np.uint32), (K, 1))
    dc, dk = torch.as_tensor(ctr.view(np.int32), device="cuda"), torch.as_tensor(key.view(np.int32), device="cuda")
    out = torch.empty_like(dc)
    _lib.check(lib.rmn_philox_raw(K, _lib.ptr(dc), _lib.ptr(dk), _lib.ptr(out), _lib.stream_ptr()))
    return out.cpu().numpy().view(np.uint32)
a, b = raw(0xFFFFFFFE), raw(0xFFFFFFFD)
u = lambda x: (x.astype(np.float64) + 0.5) / 2**32
tape = np.zeros((1, K, _lib.CP_NSLOT))
tape[0, :, 0], tape[0, :, 1], tape[0, :, 2], tape[0, :, 3] = u(a[:, 0]), u(a[:, 1]), u(a[:, 2]), u(a[:, 3])
tape[0, :, 4] = pm.xmin + (pm.xmax - pm.xmin) * u(b[:, 0])
tape[0, :, 5] = -0.1 + 0.2 * u(b[:, 1])
tape[0, :, 6] = np.floor(u(b[:, 2]) * z["k"])
tape[0, :, 7] = u(b[:, 3])
tape[0, :, 8:] = nz.cpu().numpy()
states = [unpack_state(z["k"][i], z["cpx"][i], z["cpv"][i], z["sig"][i]) for i in range(K)]
s = Sampler(m, p, states, seed=seed)
ex = s.run_injected(tape=tape); torch.cuda.synchronize()
kern = os.environ.get("RMN_CP_KERNEL", "tpc")
np.savez("/root/repo/gpurun_out/inj811_%s.npz" % kern, pk=ex["prop_k"], pcpx=ex["prop_cpx"], pcpv=ex["prop_cpv"], psig=ex["prop_sig"],
         plp=ex["prop_logpost"], lqr=ex["logqratio"], acc=ex["accepted"], k=s._chain_thetas.k, cpx=s._chain_thetas.cpx)
print("injected step ok", kern, ex["accepted"].sum())
