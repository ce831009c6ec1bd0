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
