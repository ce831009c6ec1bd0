"""Multi-GPU plumbing: one process per GPU, envs sharded in contiguous ranges, motion tables replicated.

The hot path has exactly one exchange step: the observation-normaliser moments
``[n, sum x (934), sum x^2 (934)]`` plus a few episode metrics, all fp64, packed into ONE buffer and all-reduced
(SUM) once per rollout -- a ~15 KB, latency-bound message over NVLink/NVSwitch (NCCL) -- nothing else crosses ranks.
On a CPU-only box the same code runs over gloo (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) owned by ``rank``; the first ``total % world`` ranks get one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, extra = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment; returns (rank, local_rank, world_size).
    Single-process runs (WORLD_SIZE unset or 1) do not create a process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


def pack_buffers(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])


def allreduce_packed(tensors: Sequence[torch.Tensor], group=None) -> None:
    """SUM-all-reduce several small fp64 tensors as one message, in place."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    if len(tensors) == 1 and tensors[0].dtype == torch.float64 and tensors[0].is_contiguous():
        dist.all_reduce(tensors[0], op=dist.ReduceOp.SUM, group=group)
        return
    flat = pack_buffers(tensors)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
        off += n


class EpisodeMetrics:
    """Device-resident fp64 accumulators of the per-step metrics the reference logs
    (reference puffer_phc/clean_pufferl/env.py:102-164: mean reward_raw components, resets, terminations)."""

    FIELDS = ("steps", "reward", "r_pos", "r_rot", "r_vel", "r_ang_vel", "r_power", "resets", "terminations")

    def __init__(self, device):
        self.buf = torch.zeros(len(self.FIELDS), dtype=torch.float64, device=device)

    @torch.no_grad()
    def add(self, reward, reward_raw, reset, terminated) -> None:
        n = reward.shape[0]
        self.buf[0] += n
        self.buf[1] += reward.sum(dtype=torch.float64)
        raw = reward_raw.sum(0, dtype=torch.float64)
        self.buf[2:2 + raw.shape[0]] += raw
        self.buf[7] += reset.sum(dtype=torch.float64)
        self.buf[8] += terminated.sum(dtype=torch.float64)

    def means(self) -> dict:
        b = self.buf.cpu()
        n = max(float(b[0]), 1.0)
        return {k: float(b[i]) / n for i, k in enumerate(self.FIELDS) if i > 0}


class StatsExchange:
    """The hot path's one exchange step as ONE kernel over NVLink peer memory (``phc_stats_allreduce_finalize``, csrc/stats_comm.cu):
    fold the rank's per-CTA partial sums, all-reduce ``[n, sum x, sum x^2 | episode metrics]`` across the ranks, apply
    ``RunningNorm``'s running-average update -- instead of ``phc_stats_reduce`` + ``ncclAllReduce`` + ``phc_rms_finalize`` + a memset.

    The exchange buffers are ``torch.distributed._symmetric_memory`` allocations (every rank can address every rank's buffer; the
    stores travel over NVLink / NVSwitch).  All ranks perform the same fp64 additions in the same order, so running_mean /
    running_var / count / metric sums are bit-identical across ranks.  With one rank the same kernel runs on a local buffer.
    ``StatsExchange.create`` returns ``None`` when peer memory cannot be set up (the caller then keeps the NCCL path of
    ``RunningNorm.finalize``) and says why on stderr."""

    def __init__(self, device, group=None, columns: int = 934):
        import ctypes as C
        from . import _ffi
        self.lib = _ffi.load()
        self.device = torch.device(device)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.columns = int(columns)
        n = int(self.lib.phc_stats_comm_bytes(self.world, self.columns)) // 8 + 1
        if self.world == 1:
            self.buf = torch.zeros(n, dtype=torch.float64, device=self.device)
            ptrs = [self.buf.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(n, dtype=torch.float64, device=self.device)
            self.buf.zero_()
            self._hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            torch.cuda.synchronize(self.device)
            dist.barrier(group)                      # every rank's buffer is zeroed before anyone publishes into it
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = 0
        self._comm = _ffi.StatsComm(self.rank, self.world, (C.c_void_p * 32)(*ptrs), 0, self.ticket.data_ptr())

    @classmethod
    def create(cls, device, group=None, columns: int = 934):
        try:
            return cls(device, group, columns)
        except Exception as exc:       # no NVLink peer mapping (e.g. a PCIe-only box, an old driver): the NCCL all-reduce stays
            import sys
            print(f"[puffer_phc_b200] StatsExchange unavailable ({type(exc).__name__}: {exc}); using the NCCL all-reduce", file=sys.stderr)
            return None

    @torch.no_grad()
    def allreduce_finalize(self, fused, rms=None) -> None:
        """``fused.flush_moments()`` + ``rms.finalize()`` in one launch (all ranks must call it the same number of times)."""
        import ctypes as C
        from . import _ffi
        self.epoch += 1
        self._comm.epoch = self.epoch
        mp = fused.partials if fused.accumulate_moments else None
        rows = fused._pending_rows if mp is not None else 0
        rms = rms if rms is not None else fused.rms
        upd = rms is not None and mp is not None
        with _ffi.on_device(self.device):
            _ffi.check(self.lib.phc_stats_allreduce_finalize(
                _ffi.ptr(mp), 0 if mp is None else fused.active_partials(), self.columns, int(rows), _ffi.ptr(fused.row_adjust) if mp is not None else None,
                _ffi.ptr(fused.metric_partials), fused.num_partials if fused.metrics else 0, _ffi.ptr(fused.stats), C.byref(self._comm),
                _ffi.ptr(rms.running_mean) if upd else None, _ffi.ptr(rms.running_var) if upd else None, _ffi.ptr(rms.count) if upd else None,
                _ffi.stream_ptr()), "phc_stats_allreduce_finalize")
        fused._pending_rows = 0
