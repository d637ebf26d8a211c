"""
TEST INFRASTRUCTURE -- needs the reference tree: /root/reference in the build container, or the unmodified copy
staged by oracle/build_ref.py under oracle/_ref/ (which is what travels to the GPU box).

Import the UNMODIFIED reference (rscalzo/riemann) with the harness-side shims
listed in SURVEY.md section 8c.  Nothing in the reference tree is touched; every
shim lives in ``sys.modules`` of the calling process.

Shims
-----
1. stub ``matplotlib.pyplot``      (riemann/proposals/hamiltonian.py:8 imports it
                                    for a debug plot only)
2. stub ``autograd``               (riemann/models/changepoint.py:12-13) --
                                    ``autograd.numpy`` is numpy; ``jacobian`` returns
                                    the exact analytic 2x2 Jacobian of the two
                                    mappings at changepoint.py:48-70, so the
                                    reference's own ``log|det J|`` lines (:72-78) run
3. ``sys.modules['riemann.riemann'] = riemann``   (changepoint.py:15 bad import)
4. ``riemann.proposals.MetropolisRandomWalk``     (examples/test_changepoint.py:15)
6. numpy >= 1.24 removed ``np.float`` (riemann/proposals/adaptive.py:82): ``np.float = float`` while
   the AdaptCov proposals run (value-preserving: it was an alias of the builtin).
5. numpy >= 1.24: a ``sig`` move turns ``theta.sig`` into a shape-(1,) array and
   ``model.py:50`` then builds a ragged list -> subclass that squeezes the two
   terms to floats (value-preserving).
"""
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py (unmodified copy)


def _default_root():
    """The live reference tree in the build container; on the GPU box the unmodified copy that
    `oracle/build_ref.py` staged under oracle/_ref/ (git-ignored, travels with the snapshot like the .so)."""
    env = os.environ.get("RIEMANN_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/riemann"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _default_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "riemann"))


def _analytic_jacobian(fn):
    """Stand-in for autograd.jacobian for the two maps in changepoint.py:48-70."""
    name = getattr(fn, "__name__", "")

    def jac_fwd(p):
        # (h, u) -> (h/f, h*f), f = sqrt((1-u)/u)            changepoint.py:48-59
        h, u = float(p[0]), float(p[1])
        r = (1.0 - u) / u
        f = np.sqrt(r)
        dr_du = -1.0 / (u * u)
        df_du = 0.5 / f * dr_du
        return np.array([[1.0 / f, -h / (f * f) * df_du],
                         [f, h * df_du]])

    def jac_inv(p):
        # (h1, h2) -> (sqrt(h1 h2), 1/(1+h2/h1))              changepoint.py:61-70
        h1, h2 = float(p[0]), float(p[1])
        h = np.sqrt(h1 * h2)
        q = 1.0 + h2 / h1
        return np.array([[0.5 * h2 / h, 0.5 * h1 / h],
                         [(h2 / (h1 * h1)) / (q * q), (-1.0 / h1) / (q * q)]])

    if name == "add_changepoint_mapping":
        return jac_fwd
    if name == "add_changepoint_mapping_inv":
        return jac_inv
    raise NotImplementedError("autograd shim: no analytic jacobian for %r" % name)


def _install_stubs():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if "autograd" not in sys.modules:
        try:
            import autograd  # noqa: F401
        except Exception:
            ag = types.ModuleType("autograd")
            ag.numpy = np
            ag.jacobian = _analytic_jacobian
            sys.modules["autograd"] = ag
            sys.modules["autograd.numpy"] = np
    if "emcee" not in sys.modules:
        try:
            import emcee  # noqa: F401
        except Exception:
            sys.modules["emcee"] = types.ModuleType("emcee")


_cached = None


def load_reference():
    """Return a namespace of the reference's classes (unmodified code)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    if not hasattr(np, "float"):
        np.float = float                                           # shim 6 (adaptive.py:82)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import riemann
    sys.modules["riemann.riemann"] = riemann                      # shim 3
    from riemann.samplers.sampler import Sampler
    from riemann.proposals import randomwalk, adaptive, hamiltonian
    import riemann.proposals as rprops
    rprops.MetropolisRandomWalk = randomwalk.MetropolisRandomWalk  # shim 4
    from riemann.models import gaussian, benchmarks, changepoint, model
    from riemann.samplers import ptsampler

    # examples/test_changepoint.py holds the proposal config 2 uses; load by path
    spec = importlib.util.spec_from_file_location(
        "_ref_examples_test_changepoint",
        os.path.join(REFERENCE_ROOT, "examples", "test_changepoint.py"))
    ex_cp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex_cp)

    class ChangepointRegression1DSqueezed(changepoint.ChangepointRegression1D):
        """shim 5: value-preserving float() of the two terms (numpy >= 1.24)."""

        def log_prior(self, theta):
            return float(np.squeeze(
                changepoint.ChangepointRegression1D.log_prior(self, theta)))

        def log_likelihood(self, theta):
            return float(np.squeeze(
                changepoint.ChangepointRegression1D.log_likelihood(self, theta)))

    ns = types.SimpleNamespace(
        riemann=riemann, Sampler=Sampler, Model=model.Model,
        Proposal=riemann.Proposal, ParameterError=riemann.ParameterError,
        MetropolisRandomWalk=randomwalk.MetropolisRandomWalk,
        AdaptScaleRandomWalk=randomwalk.AdaptScaleRandomWalk,
        pCN=randomwalk.pCN,
        AdaptCovRandomWalk=randomwalk.AdaptCovRandomWalk,
        AdaptScaleCovRandomWalk=randomwalk.AdaptScaleCovRandomWalk, AdaptScalepCN=randomwalk.AdaptScalepCN,
        AdaptScaleCovHMC=hamiltonian.AdaptScaleCovHMC, AdaptCovHMC=hamiltonian.AdaptCovHMC,
        AdaptScaleProposal=adaptive.AdaptScaleProposal,
        VanillaHMC=hamiltonian.VanillaHMC, AdaptScaleHMC=hamiltonian.AdaptScaleHMC,
        leapfrog=hamiltonian.leapfrog,
        MultiGaussianDist=gaussian.MultiGaussianDist, benchmarks=benchmarks,
        ChangepointParams=changepoint.ChangepointParams,
        ChangepointRegression1D=ChangepointRegression1DSqueezed,
        ChangepointRegression1DProp=ex_cp.ChangepointRegression1DProp,
        changepoint=changepoint,
        PTSampler=ptsampler.PTSampler, TemperedModel=ptsampler.TemperedModel)
    _cached = ns
    return ns


@contextlib.contextmanager
def quiet():
    """Swallow the reference's prints (examples/test_changepoint.py:64,70)."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield


class RecordingRNG(object):
    """
    Record every draw the reference makes from numpy's global legacy stream.

    The reference looks ``np.random.normal/uniform/randint`` up on the module at
    call time (randomwalk.py:25, sampler.py:84, examples/test_changepoint.py:48-67),
    so swapping the module attributes records the stream without touching the code.
    """

    def __init__(self):
        self.log = []
        self._orig = None

    def __enter__(self):
        self._orig = (np.random.normal, np.random.uniform, np.random.randint)
        o_normal, o_uniform, o_randint = self._orig
        log = self.log

        def normal(*a, **k):
            v = o_normal(*a, **k)
            log.append(("normal", np.array(v, dtype=np.float64, copy=True)))
            return v

        def uniform(*a, **k):
            v = o_uniform(*a, **k)
            log.append(("uniform", np.array(v, dtype=np.float64, copy=True)))
            return v

        def randint(*a, **k):
            v = o_randint(*a, **k)
            log.append(("randint", np.array(v, dtype=np.float64, copy=True)))
            return v

        np.random.normal, np.random.uniform, np.random.randint = normal, uniform, randint
        return self

    def __exit__(self, *exc):
        np.random.normal, np.random.uniform, np.random.randint = self._orig
        return False

    def mark(self):
        return len(self.log)
