"""CPU evidence for the round-2 idea in DESIGN.md section 5 (no GPU): the chains of a warp share ONE move-type
schedule (the three block-selection uniforms of test_changepoint.py:48-54 per step) while every other draw stays
per chain.  With the numpy port: G groups of 8 chains, each group with its own shared schedule, against the same
number of fully independent chains and against the reference's own posterior summary (tests/golden/
changepoint_posterior.npz).  Reports the posterior summaries of both designs and the correlation of the chain means
WITHIN a group, which is what the pooled B/W estimators need to be ~0.

    python scripts/shared_schedule_check.py [groups=12] [steps=16000] [burn=6000]"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import riemann_port as port      # noqa: E402   (a CPU experiment on the oracle itself)


class GroupDraws(object):
    """Draw source of one chain: the selection uniforms come from the group's shared tape, all else from its own RNG."""

    def __init__(self, sel, rng):
        self.sel, self.rng, self.t = sel, rng, 0

    def normal(self, role, n):
        return self.rng.standard_normal(n)

    def uniform(self, role, low=0.0, high=1.0):
        if role in ("sel1", "sel2", "sel3") and self.sel is not None:
            return self.sel[self.t, int(role[3]) - 1]
        return low + (high - low) * self.rng.uniform()

    def randint(self, role, n):
        return int(self.rng.integers(n))

    def next_step(self):
        self.t += 1


def run_group(args):
    gid, shared, T, burn = args
    pm, pprop, th0, _ = port.make_changepoint_problem()
    sel = np.random.default_rng(10_000 + gid).uniform(size=(T, 3)) if shared else None
    out = []
    for c in range(8):
        prop = port.ChangepointRegression1DProp(pm, pprop.hscale)
        s = port.Sampler(pm, prop, th0, draws=GroupDraws(sel, np.random.default_rng(1_000_000 * (1 + shared) + 8 * gid + c)))
        ks, sg = np.empty(T), np.empty(T)
        with np.errstate(all="ignore"):
            for t in range(T):
                th, _ = s.sample()
                ks[t], sg[t] = len(th.cpx), float(np.squeeze(th.sig))
        out.append((ks[burn:].mean(), sg[burn:].mean(), np.bincount(ks[burn:].astype(int), minlength=16)[:16] / (T - burn)))
    return out


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
    burn = int(sys.argv[3]) if len(sys.argv) > 3 else 6000
    ref = np.load(os.path.join(ROOT, "tests", "golden", "changepoint_posterior.npz"))
    with ProcessPoolExecutor(max_workers=min(os.cpu_count() or 1, 2 * G)) as ex:
        res = list(ex.map(run_group, [(g, sh, T, burn) for sh in (0, 1) for g in range(G)]))
    print("reference sampler (48 chains, fixture): mean k %.3f +- %.3f, mean sigma %.4f +- %.4f"
          % (ref["mean_k"].mean(), ref["mean_k"].std() / np.sqrt(48), ref["mean_sig"].mean(), ref["mean_sig"].std() / np.sqrt(48)))
    for sh, name in ((0, "independent schedules"), (1, "schedule shared by groups of 8")):
        grp = res[sh * G:(sh + 1) * G]
        mk = np.array([[c[0] for c in g] for g in grp])                # [G][8] chain means of k
        ms = np.array([[c[1] for c in g] for g in grp])
        n = mk.size
        # intra-group correlation of the chain means: one-way ANOVA estimate (between-group excess variance)
        def icc(m):
            msb = 8 * m.mean(axis=1).var(ddof=1)
            msw = m.var(axis=1, ddof=1).mean()
            return (msb - msw) / (msb + 7 * msw)
        print("%-32s %3d chains: mean k %.3f +- %.3f, mean sigma %.4f +- %.4f, intra-group correlation of chain means: k %+.3f, sigma %+.3f"
              % (name, n, mk.mean(), mk.std() / np.sqrt(n), ms.mean(), ms.std() / np.sqrt(n), icc(mk), icc(ms)))


if __name__ == "__main__":
    main()
