import numpy as np, sys, os, torch
sys.path.insert(0, "/root/repo")
from oracle import riemann_port as port
from riemann_b200 import Sampler, _lib
from riemann_b200.models.changepoint import ChangepointParams, ChangepointRegression1D
from riemann_b200.proposals.changepoint import ChangepointRegression1DProp
pm, pp, th0, _ = port.make_changepoint_problem()
m = ChangepointRegression1D(pm.x, pm.y, pm.xmin, pm.xmax, pm.lamb, pm.kmax, pm.alpha, pm.beta)
p = ChangepointRegression1DProp(m, pp.hscale)
K, T = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(5)
tape = np.zeros((T, K, _lib.CP_NSLOT))
tape[:, :, :4] = rng.uniform(size=(T, K, 4))
tape[:, :, 4] = rng.uniform(pm.xmin, pm.xmax, size=(T, K))
tape[:, :, 5] = rng.uniform(-0.1, 0.1, size=(T, K))
tape[:, :, 6] = rng.integers(0, 16, size=(T, K))
tape[:, :, 7] = rng.uniform(size=(T, K))
tape[:, :, 8:] = rng.standard_normal((T, K, 16))
s = Sampler(m, p, ChangepointParams([2.0], [1.0, 3.0], 0.1), K=K, seed=1)
ex = s.run_injected(tape=tape)
torch.cuda.synchronize()
tr = s._chain_thetas
np.savez("/root/repo/gpurun_out/cp_%s.npz" % os.environ.get("RMN_CP_KERNEL", "tpc"), k=tr.k, cpx=tr.cpx, cpv=tr.cpv, sig=tr.sig,
         lp=s._chain_logpost, plp=ex["prop_logpost"], acc=ex["accepted"], pk=ex["prop_k"])
print("done", os.environ.get("RMN_CP_KERNEL", "tpc"), tr.k.max(), ex["accepted"].mean())
