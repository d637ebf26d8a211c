"""Print the kernel timeline of the dense tf32x3 sampler's CUDA graph (RMN_TF32_TIMELINE=<file>; dense_tf32.cu).
usage: RMN_TF32_TIMELINE=/tmp/tl.txt python scripts/dense_timeline.py run [chains]   # runs config 3 and prints
       python scripts/dense_timeline.py show /tmp/tl.txt"""
import os
import sys


def show(path, first=6, last=6):
    replays, cur = [], None
    for line in open(path):
        if line.startswith("#"):
            cur = []
            replays.append((line.strip(), cur))
            continue
        what, h, t, a, b = line.split()
        cur.append((what, int(h[2:]), int(t[2:]), int(a), int(b)))
    hdr, ev = replays[-1]
    t0 = min(e[3] for e in ev)
    t1 = max(e[4] for e in ev)
    steps = max(e[2] for e in ev)
    print(hdr, "span %.1f us, %d steps -> %.1f us per step" % ((t1 - t0) / 1e3, steps, (t1 - t0) / 1e3 / max(steps, 1)))
    ev.sort(key=lambda e: e[3])
    mid = [e for e in ev if 5 <= e[2] <= 8]
    print("%-6s %3s %3s %10s %10s %8s" % ("kernel", "h", "t", "start us", "end us", "dur us"))
    for e in mid:
        print("%-6s %3d %3d %10.1f %10.1f %8.1f" % (e[0], e[1], e[2], (e[3] - t0) / 1e3, (e[4] - t0) / 1e3, (e[4] - e[3]) / 1e3))
    for what in ("gemm", "pass"):
        for h in sorted(set(e[1] for e in ev)):
            d = [(e[4] - e[3]) / 1e3 for e in ev if e[0] == what and e[1] == h and 2 <= e[2] < steps]
            if d:
                d.sort()
                print("%s h=%d: median %.1f us (min %.1f, max %.1f) over %d launches" % (what, h, d[len(d) // 2], d[0], d[-1], len(d)))


def run(chains):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import numpy as np
    import torch
    from riemann_b200 import Sampler
    from riemann_b200.models import benchmarks
    from riemann_b200.proposals.hamiltonian import MALA
    model = benchmarks.gauss_corr(1000)
    rng = np.random.default_rng(0)
    th0 = rng.standard_normal((chains, 1000))
    s = Sampler(model, MALA(0.08, model.grad_log_likelihood), th0, seed=1, precision="tf32x3")
    for T in (19, 19, 20):            # warm-up with another graph, then the replay that is printed
        s.run(T, trace=False)
    torch.cuda.synchronize()


if __name__ == "__main__":
    if sys.argv[1] == "run":
        path = os.environ["RMN_TF32_TIMELINE"]
        if os.path.exists(path):
            os.remove(path)
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 16384)
        show(path)
    else:
        show(sys.argv[2])
