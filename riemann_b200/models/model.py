"""
Model protocol -- same surface as riemann/models/model.py:4-64.

``Model`` is the abstract base (log_likelihood / log_prior abstract, log_posterior with
the inf/nan -> -inf rule).  ``DeviceModel`` marks models that have a CUDA kernel in the
engine; only those can be sampled by riemann_b200.Sampler (no CPU fallback).  Their
scalar-theta methods evaluate ON THE DEVICE through the pointwise C-ABI entry
(rmn_model_logpost with n = 1), so host and device never disagree.
"""
import numpy as np

from .. import _lib


class Model(object):
    def __init__(self):
        pass

    def pack(self):
        pass

    def unpack(self, theta):
        pass

    def log_likelihood(self, theta):
        raise NotImplementedError("Non-overloaded abstract method!")

    def log_prior(self, theta):
        raise NotImplementedError("Non-overloaded abstract method!")

    def log_posterior(self, theta):
        # model.py:49-54: any inf or nan in either term => -inf
        logP, logL = self.log_prior(theta), self.log_likelihood(theta)
        if not (np.isfinite(logP) and np.isfinite(logL)):
            return -np.inf
        return logP + logL

    def logL(self, theta):
        return self.log_likelihood(theta)

    def logP(self, theta):
        return self.log_prior(theta)

    def __call__(self, theta):
        return self.log_posterior(theta)


class DeviceModel(Model):
    """A Model with a device kernel.  Holds the C handle and keeps device buffers alive."""

    Ndim = None

    def __init__(self):
        self._handle = None
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.load().rmn_model_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ---- batched evaluation: Theta (n, d) -> (n,) / (n, d) on the device
    def _as_device(self, Theta):
        torch = _lib.require_cuda()
        if isinstance(Theta, torch.Tensor):
            t = Theta.to(device="cuda", dtype=torch.float64)
        else:
            t = torch.as_tensor(np.asarray(Theta, dtype=np.float64), device="cuda")
        if t.dim() == 1:
            t = t[None, :]
        if t.dim() != 2 or t.shape[1] != self.Ndim:
            from ..sampling_errors import ParameterError
            raise ParameterError("theta has shape {}, model has {} parameters"
                                 .format(tuple(t.shape), self.Ndim))
        return torch, t.contiguous()

    def _eval_batch(self, Theta, which):
        torch, t = self._as_device(Theta)
        out = torch.empty(t.shape[0], dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().rmn_model_logpost(self._handle, which, t.shape[0], _lib.ptr(t),
                                                 _lib.ptr(out), _lib.stream_ptr()))
        return out

    def log_posterior_batch(self, Theta):
        return self._eval_batch(Theta, 0)

    def log_likelihood_batch(self, Theta):
        return self._eval_batch(Theta, 1)

    def log_prior_batch(self, Theta):
        return self._eval_batch(Theta, 2)

    def grad_log_posterior_batch(self, Theta):
        torch, t = self._as_device(Theta)
        out = torch.empty_like(t)
        _lib.check(_lib.load().rmn_model_grad(self._handle, t.shape[0], _lib.ptr(t), _lib.ptr(out),
                                              _lib.stream_ptr()))
        return out

    # ---- scalar-theta protocol (model.py:27-64), evaluated on the device
    def log_likelihood(self, theta):
        return float(self._eval_batch(np.atleast_1d(theta), 1)[0])

    def log_prior(self, theta):
        return float(self._eval_batch(np.atleast_1d(theta), 2)[0])

    def log_posterior(self, theta):
        return float(self._eval_batch(np.atleast_1d(theta), 0)[0])

    def grad_log_posterior(self, theta):
        return self.grad_log_posterior_batch(np.atleast_1d(theta))[0].cpu().numpy()


def grad(fun):
    """Stand-in for ``autograd.grad`` at the one place the reference's scripts use it on a model:
    ``gradlogpost = grad(model.log_posterior)`` (examples/riemann_ex1.py:237).  For a method of a device
    model it returns that model's own gradient method -- the kernel the engine evaluates in-flight -- so
    ``VanillaHMC(eps, Nsteps, grad(model.log_posterior))`` binds to the device gradient.  Anything else has
    no device kernel and is refused (there is no automatic differentiation and no CPU fallback)."""
    from ..sampling_errors import ParameterError
    owner = getattr(fun, "__self__", None)
    name = getattr(fun, "__name__", "")
    if isinstance(owner, DeviceModel) and name in ("log_posterior", "__call__"):
        return owner.grad_log_posterior
    if isinstance(owner, DeviceModel) and name in ("log_likelihood", "logL") and hasattr(owner, "grad_log_likelihood"):
        return owner.grad_log_likelihood
    raise ParameterError("grad() takes log_posterior / log_likelihood of a riemann_b200 device model; "
                         "no device kernel exists for the gradient of an arbitrary Python callable")
