"""
Parallel tempering -- riemann/samplers/ptsampler.py:11-127 ("next" row N3 of SURVEY.md section 8f).

The reference builds one `Sampler` per temperature around `TemperedModel(model, beta)` (likelihood * beta,
prior untouched) and, every step and sequentially along the ladder, lets each chain either take a within-chain
MH step or -- with probability `Pswap` -- propose a swap with the next chain down the ladder.  Here the ladder
lies along the chain axis of the device engine: chains l*Nt .. l*Nt+Nt-1 are ladder l, the swap is a warp shuffle
inside the MH kernel on the small-d path (riemann_b200/csrc/small_gauss.cu) and an exchange of the two chain rows by the
initiator's warp on the dense-Gaussian path (d > 8, riemann_b200/csrc/dense.cu) and on the logistic model
(riemann_b200/csrc/logistic.cu; the prior is not tempered, :33-37) -- `rmn_sampler_set_tempering` -- and
`K` independent ladders run side by side.

Reference behaviours kept: the default ladder `0.5**arange(5)`; the selection uniform is drawn for every chain,
swapped or not (:104); the last chain never initiates (:107); `mhratio = min(1, exp(...))` with Python's `min`
(nan -> 1); `run(Nsamples, Nburn, Nthin)` ignores `Nburn`/`Nthin` and exposes the history of the beta = betas[0]
chain as `_chain_thetas` / `_chain_logpost` (:83-90).  Reference defects not reproduced: passing `betas=` raises
in the reference (`isinstance(betas, np.array)`, :62) -- here an explicit ladder is used as given, which is what
that branch intends; `sample()` returns `np.array(zip(...))` there (:127) and the states here.
The reference shares ONE proposal object between all temperatures (:81), so only non-adaptive proposals make
sense; the device engine enforces that (MetropolisRandomWalk, pCN; Gaussian models of any dimension, logistic model).
"""
import numpy as np

from ..models.model import Model
from ..sampling_errors import ParameterError
from .sampler import Sampler


class TemperedModel(Model):
    """ptsampler.py:11-38 -- host-side protocol object (the device kernel applies beta itself)."""

    def __init__(self, base_model, beta):
        if not (beta >= 0 and beta <= 1):
            raise ParameterError("beta = {} must be a number between 0 and 1".format(beta))
        self.base_model = base_model
        self.beta = beta

    def set_beta(self, beta):
        self.beta = beta

    def log_likelihood(self, theta):
        return (self.base_model.log_likelihood(theta)) * self.beta

    def log_prior(self, theta):
        return self.base_model.log_prior(theta)


class _TemperatureView(object):
    """`pt.samplers[i]`: the history of temperature i (what the reference keeps in its i-th Sampler)."""

    def __init__(self, pt, i):
        self._pt, self._i = pt, i
        self.model = TemperedModel(pt.model, pt.betas[i])

    @property
    def _chain_thetas(self):
        return self._pt._history_of(self._i)[0]

    @property
    def _chain_logpost(self):
        return self._pt._history_of(self._i)[1]

    def current_state(self):
        th, lp = self._pt._history_of(self._i)
        return th[-1], lp[-1]


class PTSampler(object):
    """Same constructor as the reference (ptsampler.py:47) plus `K` = number of independent ladders."""

    def __init__(self, model, proposal, theta0, betas=None, Pswap=0.1, K=1, seed=0, chain_offset=0):
        self.betas = 0.5 ** np.arange(5) if betas is None else np.asarray(betas, dtype=np.float64)
        if self.betas.ndim != 1 or not (2 <= len(self.betas) <= 32):
            raise ParameterError("betas must be a 1-d ladder of 2..32 temperatures")
        if not (Pswap > 0 and Pswap < 1):
            raise ParameterError("Pswap must be a number between 0 and 1")
        self.Pswap, self.model, self.proposal = Pswap, model, proposal
        self.K, self.Nt = int(K), len(self.betas)
        th0 = np.atleast_1d(np.asarray(theta0, dtype=np.float64))
        if th0.ndim == 1:
            th0 = np.tile(th0, (self.K * self.Nt, 1))                  # every temperature starts at theta0 (:80-81)
        elif th0.shape[0] == self.K:
            th0 = np.repeat(th0, self.Nt, axis=0)
        elif th0.shape[0] != self.K * self.Nt:
            raise ParameterError("theta0 must be (d,), (K, d) or (K*Nt, d)")
        self._sampler = Sampler(model, proposal, th0, seed=seed, chain_offset=chain_offset * self.Nt,
                                _tempering=(self.betas, Pswap))
        self.d = self._sampler.d
        self._thetas = [np.asarray(self._sampler._chain_thetas[0]).reshape(self.K, self.Nt, self.d)]
        self._logpost = [np.asarray(self._sampler._chain_logpost[0]).reshape(self.K, self.Nt)]
        self.samplers = [_TemperatureView(self, i) for i in range(self.Nt)]
        self._publish()

    # history [record][ladder][temperature]
    def _history_of(self, i):
        th = np.stack([t[:, i] for t in self._thetas])                 # [rec, K, d]
        lp = np.stack([l[:, i] for l in self._logpost])
        if self.K == 1:
            return list(th[:, 0]), list(lp[:, 0])
        return th, lp

    def _publish(self):
        self._chain_thetas, self._chain_logpost = self._history_of(0)

    def _absorb(self, ex=None):
        th = np.asarray(self._sampler._chain_thetas)
        lp = np.asarray(self._sampler._chain_logpost)
        if th.ndim == 2:                                               # a single chain comes back as [rec, d]
            th, lp = th[:, None, :], lp[:, None]
        for r in range(1, th.shape[0]):                                # record 0 is the state we already hold
            self._thetas.append(th[r].reshape(self.K, self.Nt, self.d))
            self._logpost.append(lp[r].reshape(self.K, self.Nt))
        self._publish()
        return ex

    def run(self, Nsamples, Nburn=0, Nthin=1, trace=True):
        """ptsampler.py:83-90: Nsamples PT steps; Nburn / Nthin are accepted and ignored like there.
        `trace=False` keeps only the final state (many ladders, long runs)."""
        if trace:
            self._sampler.run(int(Nsamples), 0, 1)
            self._absorb()
        else:
            self._sampler.run(int(Nsamples), trace=False)
            self._thetas = [np.asarray(self._sampler._chain_thetas[-1]).reshape(self.K, self.Nt, self.d)]
            self._logpost = [np.asarray(self._sampler._chain_logpost[-1]).reshape(self.K, self.Nt)]
            self._publish()

    def run_injected(self, usel, xi, u):
        """Parity mode: usel[T, K*Nt], xi[T, K*Nt, d], u[T, K*Nt] in the reference's per-chain draw order."""
        ex = self._sampler.run_injected(xi=xi, u=u, usel=usel)
        return self._absorb(ex)

    def sample(self):
        self.run(1)
        return [s.current_state() for s in self.samplers]

    def diagnostics(self, allreduce=True):
        """Posterior summaries of the beta = betas[0] chains ONLY (one per ladder): the device accumulators hold every
        rung of every ladder, and pooling the tempered rungs (beta < 1) would inflate the variance.  Built from the
        per-chain moments (`Sampler.chain_moments`), rank-local (`allreduce` is accepted for interface symmetry:
        ladders on other ranks are independent replicas, combine their summaries on the host if needed).
        `accept_rate` is that of the flat engine -- accepted within-chain moves per chain-step over all temperatures,
        a swap step counting as no move -- and is reported as `accept_rate_all_rungs`; `swap_fraction` = Pswap."""
        from ..distributed import summarize_block
        flat = self._sampler.diagnostics(allreduce=False)
        m, v = self._sampler.chain_moments()                      # [nd, K * Nt], chain fastest
        m0, v0 = m[:, ::self.Nt], v[:, ::self.Nt]                 # rung 0 of every ladder
        nd = m0.shape[0]
        blk = np.zeros(6 + 3 * nd)
        blk[0], blk[1], blk[4] = self.K, flat["samples"], flat["steps"]
        blk[6:6 + nd] = m0.sum(axis=1)
        blk[6 + nd:6 + 2 * nd] = (m0 * m0).sum(axis=1)
        blk[6 + 2 * nd:] = v0.sum(axis=1)
        out = summarize_block(blk)
        out.pop("accept_rate", None)
        out["accept_rate_all_rungs"] = flat["accept_rate"]
        out["swap_fraction"] = float(self.Pswap)
        out["temperatures"] = self.Nt
        return out
