"""Device-resident rollout buffer ("next" row f2 of SURVEY.md section 8): the part of the reference's ``Experience``
(reference puffer_phc/clean_pufferl/structs.py:23-176) and of ``clean_pufferl.train`` (core.py:213-259) that sits
around ``c_gae.compute_gae``.

The reference keeps values / rewards / dones in host numpy arrays in *arrival order* (per step: the rows of the
non-masked envs, until ``batch_size`` rows are stored), sorts them by ``(env_id, step)`` with a Python ``sorted()`` over
131072 tuples, runs GAE on the host and copies the advantages back.  Here everything stays in HBM in a fixed ``[T, N]``
layout (+ mask); ``store`` is one launch per env step (``phc_rollout_store``), the reference's arrival indices and its sorted
order come out of ``phc_rollout_sort`` (four launches: prefix sums over 32-env group counts, per-env offsets, a warp-per-env
compaction that writes the sorted arrays contiguously), and the scan runs in ``phc_gae`` -- no host round trip, one sync per
rollout (the size of the ragged result).  CUDA only: there is no host implementation.

Semantics kept from the reference (all verified against a literal Python replay in tests/test_rollout.py):
* ``store`` keeps, per step, only rows with ``mask`` true (truncated envs are masked out, clean_pufferl/env.py:133) and
  stops at ``batch_size`` rows -- the last step may be stored partially (structs.py:116);
* ``sort_training_data`` returns the arrival indices ordered by (env_id, step) (structs.py:133-145);
* ``compute_advantages`` = ``compute_gae(dones[idxs], values[idxs], rewards[idxs] (+ extra), gamma, lambda)`` over that
  flat env-major sequence, with the carry crossing env boundaries exactly like the reference's flat scan (core.py:249).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _ffi
from .c_gae import compute_gae_cuda


class RolloutBuffer:
    def __init__(self, num_envs: int, batch_size: int, max_steps: Optional[int] = None, obs_dim: int = 0, device="cuda"):
        self.N, self.batch_size = int(num_envs), int(batch_size)
        # masked rows make a rollout longer than batch_size / num_envs steps; leave room (grown on demand)
        self.T = int(max_steps) if max_steps is not None else 2 * (-(-self.batch_size // self.N)) + 2
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RolloutBuffer: the rollout kernels are CUDA only (no host implementation)")
        self.lib = _ffi.load()
        f = dict(dtype=torch.float32, device=self.device)
        self.groups = -(-self.N // 32)
        self.values = torch.zeros(self.T, self.N, **f)
        self.rewards = torch.zeros(self.T, self.N, **f)
        self.dones = torch.zeros(self.T, self.N, **f)
        self.truncateds = torch.zeros(self.T, self.N, **f)
        self.mask = torch.zeros(self.T, self.N, dtype=torch.bool, device=self.device)
        self._subrank = torch.zeros(self.T, self.N, dtype=torch.uint8, device=self.device)          # non-masked envs before e in its 32-env group
        self._group_counts = torch.zeros(self.T, self.groups, dtype=torch.int32, device=self.device)
        self._row_counts = torch.zeros(self.T, dtype=torch.int32, device=self.device)
        self.obs = torch.zeros(self.T, self.N, obs_dim, **f) if obs_dim else None
        self.step = 0
        self._stored = torch.zeros(1, dtype=torch.int64, device=self.device)      # rows stored so far (device counter, no sync)
        self._meta = torch.zeros(4, dtype=torch.int64, device=self.device)
        cap = self.batch_size
        self._sorted_buf = {k: torch.empty(cap, **f) for k in ("dones", "values", "rewards")}
        self._idxs_buf = torch.empty(cap, dtype=torch.int64, device=self.device)
        self._pos_buf = torch.empty(cap, dtype=torch.int64, device=self.device)
        self._scratch = None
        self._base = None
        self._stored_ptr = _ffi.ptr(self._stored)
        self.idxs = None

    # ---- structs.py:108-131 -------------------------------------------------------------------------------
    def store(self, value, reward, done, trunc, mask, obs=None) -> None:
        """One env step for all N envs (env_id = arange(N)): ONE launch, everything stays on the device, nothing syncs."""
        t = self.step
        if t >= self.T:
            self._grow()
        f32, N = torch.float32, self.N

        def vec(x, dt):                            # the usual case (right dtype, contiguous) costs two attribute reads
            if not x.is_cuda:
                raise RuntimeError("puffer_phc_b200: expected CUDA tensors -- the kernels are the only implementation "
                                   f"(got a tensor on {x.device})")
            if x.dtype is not dt or not x.is_contiguous():
                x = x.to(dt).contiguous()
            if x.numel() != N:
                raise ValueError("RolloutBuffer.store: every per-env vector must have num_envs elements")
            return x
        value, reward = vec(value, f32), vec(reward, f32)
        flags_float = done.dtype is not torch.bool and done.dtype is not torch.uint8
        fdt = f32 if flags_float else done.dtype
        done = vec(done, fdt)
        trunc = None if trunc is None else vec(trunc, fdt)
        mask = vec(mask, torch.bool)
        if self._base is None:                     # row pointers are base + t * row bytes: no tensor slicing per call
            self._base = tuple(x.data_ptr() for x in (self.values, self.rewards, self.dones, self.truncateds, self.mask, self._subrank,
                                                      self._group_counts, self._row_counts))
        bv, br, bd, bt, bm, bs, bg, bc = self._base
        r4, r1 = 4 * t * N, t * N
        P = _ffi.C.c_void_p
        with _ffi.on_device(self.device):
            _ffi.check(self.lib.phc_rollout_store(P(value.data_ptr()), P(reward.data_ptr()), P(done.data_ptr()),
                                                  P(None if trunc is None else trunc.data_ptr()), int(flags_float), P(mask.data_ptr()), N,
                                                  P(bv + r4), P(br + r4), P(bd + r4), P(bt + r4), P(bm + r1), P(bs + r1),
                                                  P(bg + 4 * t * self.groups), P(bc + 4 * t), self._stored_ptr, _ffi.stream_ptr()),
                       "RolloutBuffer.store")
        if self.obs is not None and obs is not None:
            self.obs[t].copy_(obs)
        self.step += 1
        self.idxs = None

    def _grow(self) -> None:
        for name in ("values", "rewards", "dones", "truncateds", "mask", "obs", "_subrank", "_group_counts", "_row_counts"):
            a = getattr(self, name)
            if a is not None:
                setattr(self, name, torch.cat([a, torch.zeros_like(a)], 0))
        self.T *= 2
        self._scratch = None
        self._base = None

    @property
    def full(self) -> bool:                       # structs.py:104-106 (one device->host read)
        return int(self._stored) >= self.batch_size

    # ---- structs.py:133-145 -------------------------------------------------------------------------------
    def sort_training_data(self) -> torch.Tensor:
        """Arrival indices (the reference's row numbers) in (env_id, step) order, as a device int64 tensor.  Also leaves dones /
        values / rewards in that order for ``compute_advantages``."""
        T = self.step
        if T == 0:
            raise RuntimeError("RolloutBuffer.sort_training_data: nothing stored")
        need = int(self.lib.phc_rollout_scratch_bytes(self.N, self.T))
        if self._scratch is None or self._scratch.numel() * 8 < need:
            self._scratch = torch.empty((need + 7) // 8, dtype=torch.int64, device=self.device)
        sb = self._sorted_buf
        with _ffi.on_device(self.device):
            _ffi.check(self.lib.phc_rollout_sort(_ffi.ptr(self.dones), _ffi.ptr(self.values), _ffi.ptr(self.rewards), _ffi.ptr(self.mask),
                                                 _ffi.ptr(self._subrank), _ffi.ptr(self._group_counts), _ffi.ptr(self._row_counts), self.N, T,
                                                 self.batch_size, _ffi.ptr(self._scratch), _ffi.ptr(self._meta), _ffi.ptr(sb["dones"]),
                                                 _ffi.ptr(sb["values"]), _ffi.ptr(sb["rewards"]), _ffi.ptr(self._idxs_buf),
                                                 _ffi.ptr(self._pos_buf), _ffi.stream_ptr()), "RolloutBuffer.sort_training_data")
        rows = int(self._meta[2])                                                     # the one sync of the rollout
        self.idxs = self._idxs_buf[:rows]
        self._pos_em = self._pos_buf[:rows]                                           # positions in the env-major [N*T] flattening
        self._rows = rows
        return self.idxs

    def _sorted(self, a: torch.Tensor) -> torch.Tensor:
        return a[: self.step].t().reshape(-1)[self._pos_em]

    # ---- core.py:213-259 ----------------------------------------------------------------------------------
    def compute_advantages(self, gamma: float, gae_lambda: float, extra_reward: Optional[torch.Tensor] = None):
        """Advantages and returns in the reference's sorted order (length = number of stored rows)."""
        if self.idxs is None:
            self.sort_training_data()
        n = self._rows
        d, v, r = (self._sorted_buf[k][:n] for k in ("dones", "values", "rewards"))
        if extra_reward is not None:
            r = r + extra_reward.reshape(-1)                                          # adversarial reward, core.py:249
        adv = compute_gae_cuda(d, v, r, gamma, gae_lambda)
        self.advantages, self.sorted_values = adv, v
        self.returns = adv + v                                                        # b_returns = b_advantages + b_values
        return adv, self.returns

    def sorted_obs(self) -> torch.Tensor:
        """Observation rows in the same sorted order (what ``obs[b_idxs_obs]`` gathers, structs.py:155)."""
        T = self.step
        return self.obs[:T].transpose(0, 1).reshape(self.N * T, -1)[self._pos_em]

    def reset(self) -> None:
        self.step = 0
        self._stored.zero_()
        self._row_counts.zero_()
        self.idxs = None
