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


# ---- the forward as a registered operator --------------------------------------------------------------------------------------
# The reference scripts this module (``torch.jit.script(RunningNorm(n))``, reference puffer_phc/policies/discriminator_policy.py:21), so
# ``forward`` must be TorchScript-compilable: it calls ``torch.ops.phc_b200.rms_forward``, a torch.library operator whose CUDA
# implementation is the C-ABI kernel (csrc/rms.cu).  There is no CPU kernel on purpose: the dispatcher raises for host tensors.
_LIB = torch.library.Library("phc_b200", "DEF")
_LIB.define("rms_forward(Tensor x, Tensor mean, Tensor var, float eps, float clip) -> Tensor")


def _rms_forward_cuda(x, mean, var, eps, clip):
    lib = _ffi.load()
    C_ = mean.shape[-1]
    x2 = x.reshape(-1, C_)
    if x2.dtype != torch.float32:
        x2 = x2.float()
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    y = torch.empty((x2.shape[0], C_), dtype=torch.float32, device=x.device)
    with _ffi.on_device(x.device):
        _ffi.check(lib.phc_rms_forward(_ffi.ptr(x2), x2.stride(0), _ffi.ptr(mean), _ffi.ptr(var), float(eps), float(clip), x2.shape[0],
                                       C_, _ffi.ptr(y), y.stride(0), _ffi.stream_ptr()), "RunningNorm.forward")
    return y.view(x.shape)


def _rms_forward_meta(x, mean, var, eps, clip):
    return torch.empty(x.shape, dtype=torch.float32, device=x.device)


_LIB.impl("rms_forward", _rms_forward_cuda, "CUDA")
_LIB.impl("rms_forward", _rms_forward_meta, "Meta")


class RunningNorm(nn.Module):
    def __init__(self, shape: int, epsilon=1e-5, clip=10.0):
        super().__init__()
        self.register_buffer("running_mean", torch.zeros((1, shape), dtype=torch.float32))
        self.register_buffer("running_var", torch.ones((1, shape), dtype=torch.float32))
        self.register_buffer("count", torch.ones(1, dtype=torch.float32))
        self.epsilon = float(epsilon)
        self.clip = float(clip)
        # state of the split update, kept as NON-persistent buffers (not in the state_dict, TorchScript-friendly, follow .to()):
        # pending moments fp64 [1 + 2C] = n, sum, sum of squares; the enclosing statistics buffer (attach_stats) and reduction scratch
        self.register_buffer("_moments", torch.zeros(1 + 2 * shape, dtype=torch.float64), persistent=False)
        self.register_buffer("_stats", torch.zeros(0, dtype=torch.float64), persistent=False)
        self.register_buffer("_scratch", torch.zeros(0, dtype=torch.float64), persistent=False)

    # ---- forward (running_norm.py:15-20) -----------------------------------------------------------
    def forward(self, x):
        return torch.ops.phc_b200.rms_forward(x, self.running_mean, self.running_var, self.epsilon, self.clip)

    # ---- update (running_norm.py:23-34), split into accumulate / (all-reduce) / finalize -----------------
    @torch.jit.ignore
    def attach_stats(self, stats: torch.Tensor) -> None:
        """Make the pending moments the head of ``stats`` = fp64 ``[n, sum x (C), sum x^2 (C), episode metrics (NUM_METRICS)]`` -- the
        rank's ONE statistics buffer (FusedStep.stats): ``finalize()`` then all-reduces moments and metrics as a single message."""
        C_ = self.running_mean.shape[1]
        assert stats.dtype == torch.float64 and stats.is_contiguous() and stats.numel() >= 1 + 2 * C_
        self._stats = stats
        self._moments = stats[: 1 + 2 * C_]

    @torch.jit.ignore
    def moments_buffer(self) -> torch.Tensor:
        if self._moments.device != self.running_mean.device:
            self._moments = torch.zeros(self._moments.shape, dtype=torch.float64, device=self.running_mean.device)
        return self._moments

    @torch.jit.ignore
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
        if self._scratch.numel() < n or self._scratch.device != x.device:
            self._scratch = torch.empty(n, dtype=torch.float64, device=x.device)
        with _ffi.on_device(x.device):
            _ffi.check(lib.phc_rms_moments(_ffi.ptr(x), x.stride(0), x.shape[0], C_, _ffi.ptr(self.moments_buffer()),
                                           _ffi.ptr(self._scratch), _ffi.stream_ptr()), "RunningNorm.accumulate")

    @torch.jit.ignore
    @torch.no_grad()
    def accumulate_partials(self, partials: torch.Tensor, rows: int) -> None:
        """Fold the per-CTA partial sums written by the fused step kernel into the pending moments."""
        lib = _ffi.load()
        C_ = self.running_mean.shape[1]
        with _ffi.on_device(partials.device):
            _ffi.check(lib.phc_rms_reduce_partials(_ffi.ptr(partials), partials.shape[0], int(rows), C_,
                                                   _ffi.ptr(self.moments_buffer()), _ffi.stream_ptr()), "RunningNorm.accumulate_partials")

    @torch.jit.ignore
    @torch.no_grad()
    def finalize(self, group=None, allreduce: bool = True) -> None:
        """All-reduce the pending moments over ``group`` (if torch.distributed is initialised; ``allreduce=False`` keeps the update
        rank-local) and apply the reference's running-average update; clears the pending moments."""
        lib = _ffi.load()
        m = self.moments_buffer()
        if allreduce and torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
            buf = self._stats if (self._stats.numel() > 0 and self._moments.data_ptr() == self._stats.data_ptr()) else m
            torch.distributed.all_reduce(buf, op=torch.distributed.ReduceOp.SUM, group=group)
        C_ = self.running_mean.shape[1]
        with _ffi.on_device(m.device):
            _ffi.check(lib.phc_rms_finalize(_ffi.ptr(m), C_, _ffi.ptr(self.running_mean), _ffi.ptr(self.running_var),
                                            _ffi.ptr(self.count), _ffi.stream_ptr()), "RunningNorm.finalize")
        m.zero_()

    @torch.jit.ignore
    @torch.no_grad()
    def update(self, x, group=None):
        """running_norm.py:23-34: one call = one equal-weight running-average step of the batch mean / biased var."""
        assert x.dim() == 2, "x must be 2D"
        self.moments_buffer().zero_()
        self.accumulate(x)
        self.finalize(group)

    # ---- pickling hooks kept from the reference (running_norm.py:37-53) ----------------------------
    @torch.jit.ignore
    def __getstate__(self):
        return {"running_mean": self.running_mean, "running_var": self.running_var, "count": self.count,
                "epsilon": self.epsilon, "clip": self.clip}

    @torch.jit.ignore
    def __setstate__(self, state):
        nn.Module.__init__(self)
        self.register_buffer("running_mean", state["running_mean"])
        self.register_buffer("running_var", state["running_var"])
        self.register_buffer("count", state["count"])
        self.epsilon = float(state["epsilon"])
        self.clip = float(state["clip"])
        C_ = state["running_mean"].shape[-1]
        dev = state["running_mean"].device
        self.register_buffer("_moments", torch.zeros(1 + 2 * C_, dtype=torch.float64, device=dev), persistent=False)
        self.register_buffer("_stats", torch.zeros(0, dtype=torch.float64, device=dev), persistent=False)
        self.register_buffer("_scratch", torch.zeros(0, dtype=torch.float64, device=dev), persistent=False)
