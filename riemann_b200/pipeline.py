"""
Host-buffer job pipeline: many independent sampling jobs (start states in pinned host memory -> T MH iterations ->
final states, log-posteriors and the diagnostics block back in pinned host memory) through ONE device sampler.

The reference has no counterpart (its `Sampler.run`, sampler.py:44-54, works on host lists); this is the host-side
loop a user of the device engine writes when the chains' start states arrive from the host job after job.  Uploads run
on a copy-in stream, downloads on a copy-out stream, both double-buffered, so the copies of job i+1 / i-1 overlap the MH
kernels of job i; the kernels themselves run on the caller's current stream, one job after the other (the sampler has
one state).  With `overlap=False` every job is upload -> run -> download -> synchronize, the plain sequence.

    runner = HostJobRunner(sampler)
    for out in runner.run(jobs, T):        # jobs: iterable of tuples of pinned host tensors (see `host_layout`)
        ...                                 # out: dict(state=[...pinned tensors...], logpost=..., diagnostics=...)

The yielded buffers are reused two jobs later: consume (or copy) them before advancing the generator twice.
"""
from . import _lib


class HostJobRunner(object):

    def __init__(self, sampler, overlap=True):
        self.s = sampler
        self.overlap = bool(overlap)
        torch = self.torch = sampler._torch
        lib = self.lib = _lib.load()
        K = sampler.K
        if sampler._is_cp:
            from .models.changepoint import LANES
            self._shapes = [((K,), torch.int32), ((K, LANES), torch.float64), ((K, LANES), torch.float64),
                            ((K,), torch.float64)]
        else:
            self._shapes = [((K, sampler.d), torch.float64)]
        nd = lib.rmn_sampler_diag_dim(sampler._handle)
        self._nblk = _lib.DIAG_HDR + 3 * nd
        nslot = 2 if self.overlap else 1
        mk = lambda dev, pin: [[self._empty(sh, dt, dev, pin) for sh, dt in self._shapes] for _ in range(nslot)]
        self.dev_in, self.dev_out, self.host_out = mk("cuda", False), mk("cuda", False), mk("cpu", True)
        self.dev_lp = [torch.empty(K, dtype=torch.float64, device="cuda") for _ in range(nslot)]
        self.host_lp = [torch.empty(K, dtype=torch.float64).pin_memory() for _ in range(nslot)]
        self.dev_blk = [torch.empty(self._nblk, dtype=torch.float64, device="cuda") for _ in range(nslot)]
        self.host_blk = [torch.empty(self._nblk, dtype=torch.float64).pin_memory() for _ in range(nslot)]
        self.s_in = torch.cuda.Stream() if self.overlap else None
        self.s_out = torch.cuda.Stream() if self.overlap else None
        ev = lambda: [torch.cuda.Event() for _ in range(nslot)]
        self.ev_uploaded, self.ev_in_free, self.ev_ready, self.ev_downloaded = ev(), ev(), ev(), ev()
        self.h2d_bytes = sum(self._nbytes(sh, dt) for sh, dt in self._shapes)
        self.d2h_bytes = self.h2d_bytes + K * 8 + self._nblk * 8

    def _empty(self, shape, dtype, dev, pin):
        t = self.torch.empty(shape, dtype=dtype, device=dev)
        return t.pin_memory() if pin else t

    def _nbytes(self, shape, dtype):
        n = 1
        for v in shape:
            n *= v
        return n * (4 if dtype == self.torch.int32 else 8)

    def host_layout(self):
        """[(shape, dtype)] of one job's input tensors: theta[K, d] (fp64), or k, cpx, cpv, sig for the changepoint model."""
        return list(self._shapes)

    # -- the three stages of one job ------------------------------------------------------------------------------
    def _upload(self, job, b, stream):
        with self.torch.cuda.stream(stream):
            for h, d in zip(job, self.dev_in[b]):
                d.copy_(h, non_blocking=True)

    def _compute(self, b, T):
        s, lib, sp = self.s, self.lib, _lib.stream_ptr()
        if s._is_cp:
            _lib.check(lib.rmn_sampler_cp_set_state(s._handle, *[_lib.ptr(d) for d in self.dev_in[b]], sp))
        else:
            _lib.check(lib.rmn_sampler_set_state(s._handle, _lib.ptr(self.dev_in[b][0]), sp))
        return sp

    def _finish(self, b, T):
        s, lib, sp = self.s, self.lib, _lib.stream_ptr()
        _lib.check(lib.rmn_sampler_run(s._handle, int(T), None, None, sp))
        if s._is_cp:
            _lib.check(lib.rmn_sampler_cp_get_state(s._handle, *[_lib.ptr(d) for d in self.dev_out[b]],
                                                    _lib.ptr(self.dev_lp[b]), sp))
        else:
            _lib.check(lib.rmn_sampler_get_state(s._handle, _lib.ptr(self.dev_out[b][0]), _lib.ptr(self.dev_lp[b]), sp))
        _lib.check(lib.rmn_sampler_reduce_diagnostics(s._handle, _lib.ptr(self.dev_blk[b]), sp))

    def _download(self, b, stream):
        with self.torch.cuda.stream(stream):
            for h, d in zip(self.host_out[b], self.dev_out[b]):
                h.copy_(d, non_blocking=True)
            self.host_lp[b].copy_(self.dev_lp[b], non_blocking=True)
            self.host_blk[b].copy_(self.dev_blk[b], non_blocking=True)

    def _result(self, b):
        return {"state": self.host_out[b], "logpost": self.host_lp[b], "diagnostics": self.host_blk[b]}

    def run(self, jobs, T):
        """Generator over the jobs' results, in order."""
        torch = self.torch
        main = torch.cuda.current_stream()
        if not self.overlap:
            for job in jobs:
                self._upload(job, 0, main)
                self._compute(0, T)
                self._finish(0, T)
                self._download(0, main)
                main.synchronize()
                yield self._result(0)
            return
        pending = []                                   # slots whose download is in flight, oldest first
        for i, job in enumerate(jobs):
            b = i & 1
            if i >= 2:
                # slot b is being reused: job i-2's download must have landed, and its result goes to the caller first
                self.ev_downloaded[b].synchronize()
                pending.pop(0)
                yield self._result(b)
                self.s_in.wait_event(self.ev_in_free[b])       # set_state of job i-2 has consumed dev_in[b]
            self._upload(job, b, self.s_in)
            self.ev_uploaded[b].record(self.s_in)
            main.wait_event(self.ev_uploaded[b])
            self._compute(b, T)
            self.ev_in_free[b].record(main)
            if i >= 2:
                main.wait_event(self.ev_downloaded[b])         # dev_out[b] has been read out (already true: synchronised)
            self._finish(b, T)
            self.ev_ready[b].record(main)
            self.s_out.wait_event(self.ev_ready[b])
            self._download(b, self.s_out)
            self.ev_downloaded[b].record(self.s_out)
            pending.append(b)
        for b in pending:
            self.ev_downloaded[b].synchronize()
            yield self._result(b)
