"""Drop-in for the reference's ``RunningNorm`` (reference puffer_phc/policies/running_norm.py:5-53): same
constructor, buffers (``running_mean [1,C]``, ``running_var [1,C]``, ``count [1]``), ``forward``/``update`` and
pickling hooks, so a reference ``state_dict`` round-trips unchanged.

``forward`` is one CUDA kernel (csrc/rms.cu).  ``update`` is split the B200 way: per-column fp64 moments
(``[n, sum x, sum x^2]``) are accumulated on device -- either by ``phc_rms_moments`` over a rollout buffer or, for
free, inside the fused step kernel -- optionally all-reduced across ranks (NCCL over NVLink, one ~15 KB message),
then ``phc_rms_finalize`` applies the reference's running-average rule.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _ffi


class RunningNorm(nn.Module):
    def __init__(self, shape: int, epsilon=1e-5, clip=10.0):
        super().__init__()
        self.register_buffer("running_mean", torch.zeros((1, shape), dtype=torch.float32))
        self.register_buffer("running_var", torch.ones((1, shape), dtype=torch.float32))
        self.register_buffer("count", torch.ones(1, dtype=torch.float32))
        self.epsilon = epsilon
        self.clip = clip
        self._moments = None      # fp64 [1 + 2C]: n, sum, sum of squares (device)
        self._stats = None        # optional enclosing buffer [1 + 2C + NUM_METRICS] (attach_stats): what finalize() all-reduces
        self._scratch = None

    # ---- forward (running_norm.py:15-20) -----------------------------------------------------------
    def forward(self, x):
        lib = _ffi.load()
        _ffi.require_cuda(x, self.running_mean)
        C_ = self.running_mean.shape[1]
        x2 = x.reshape(-1, C_)
        if x2.dtype != torch.float32:
            x2 = x2.float()
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        y = torch.empty((x2.shape[0], C_), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _ffi.check(lib.phc_rms_forward(_ffi.ptr(x2), x2.stride(0), _ffi.ptr(self.running_mean), _ffi.ptr(self.running_var),
                                           float(self.epsilon), float(self.clip), x2.shape[0], C_, _ffi.ptr(y), y.stride(0),
                                           _ffi.stream_ptr()), "RunningNorm.forward")
        return y.view(x.shape)

    # ---- update (running_norm.py:23-34), split into accumulate / (all-reduce) / finalize -----------------
    def attach_stats(self, stats: torch.Tensor) -> None:
        """Make the pending moments the head of ``stats`` = fp64 ``[n, sum x (C), sum x^2 (C), episode metrics (NUM_METRICS)]`` -- the
        rank's ONE statistics buffer (FusedStep.stats): ``finalize()`` then all-reduces moments and metrics as a single message."""
        C_ = self.running_mean.shape[1]
        assert stats.dtype == torch.float64 and stats.is_contiguous() and stats.numel() >= 1 + 2 * C_
        self._stats = stats
        self._moments = stats[: 1 + 2 * C_]

    def moments_buffer(self) -> torch.Tensor:
        C_ = self.running_mean.shape[1]
        if self._moments is None or self._moments.device != self.running_mean.device:
            self._moments = torch.zeros(1 + 2 * C_, dtype=torch.float64, device=self.running_mean.device)
        return self._moments

    @torch.no_grad()
    def accumulate(self, x) -> None:
        """Add the rows of ``x [B, C]`` to the pending moments."""
        lib = _ffi.load()
        _ffi.require_cuda(x, self.running_mean)
        C_ = self.running_mean.shape[1]
        assert x.dim() == 2 and x.shape[1] == C_, "x must be 2D [B, C]"
        x = x.float()
        if x.stride(1) != 1:
            x = x.contiguous()
        n = int(lib.phc_rms_scratch_doubles(C_))
        if self._scratch is None or self._scratch.numel() < n or self._scratch.device != x.device:
            self._scratch = torch.empty(n, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _ffi.check(lib.phc_rms_moments(_ffi.ptr(x), x.stride(0), x.shape[0], C_, _ffi.ptr(self.moments_buffer()),
                                           _ffi.ptr(self._scratch), _ffi.stream_ptr()), "RunningNorm.accumulate")

    @torch.no_grad()
    def accumulate_partials(self, partials: torch.Tensor, rows: int) -> None:
        """Fold the per-CTA partial sums written by the fused step kernel into the pending moments."""
        lib = _ffi.load()
        C_ = self.running_mean.shape[1]
        with torch.cuda.device(partials.device):
            _ffi.check(lib.phc_rms_reduce_partials(_ffi.ptr(partials), partials.shape[0], int(rows), C_,
                                                   _ffi.ptr(self.moments_buffer()), _ffi.stream_ptr()), "RunningNorm.accumulate_partials")

    @torch.no_grad()
    def finalize(self, group=None, allreduce: bool = True) -> None:
        """All-reduce the pending moments over ``group`` (if torch.distributed is initialised; ``allreduce=False`` keeps the update
        rank-local) and apply the reference's running-average update; clears the pending moments."""
        lib = _ffi.load()
        m = self.moments_buffer()
        if allreduce and torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
            buf = self._stats if (self._stats is not None and self._moments.data_ptr() == self._stats.data_ptr()) else m
            torch.distributed.all_reduce(buf, op=torch.distributed.ReduceOp.SUM, group=group)
        C_ = self.running_mean.shape[1]
        with torch.cuda.device(m.device):
            _ffi.check(lib.phc_rms_finalize(_ffi.ptr(m), C_, _ffi.ptr(self.running_mean), _ffi.ptr(self.running_var),
                                            _ffi.ptr(self.count), _ffi.stream_ptr()), "RunningNorm.finalize")
        m.zero_()

    @torch.no_grad()
    def update(self, x, group=None):
        """running_norm.py:23-34: one call = one equal-weight running-average step of the batch mean / biased var."""
        assert x.dim() == 2, "x must be 2D"
        self.moments_buffer().zero_()
        self.accumulate(x)
        self.finalize(group)

    # ---- pickling hooks kept from the reference (running_norm.py:37-53) ----------------------------
    def __getstate__(self):
        return {"running_mean": self.running_mean, "running_var": self.running_var, "count": self.count,
                "epsilon": self.epsilon, "clip": self.clip}

    def __setstate__(self, state):
        nn.Module.__init__(self)
        self.register_buffer("running_mean", state["running_mean"])
        self.register_buffer("running_var", state["running_var"])
        self.register_buffer("count", state["count"])
        self.epsilon = state["epsilon"]
        self.clip = state["clip"]
        self._moments = None
        self._stats = None
        self._scratch = None
