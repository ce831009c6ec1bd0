"""Device-resident rollout buffer ("next" row f2 of SURVEY.md section 8): the part of the reference's ``Experience``
(reference puffer_phc/clean_pufferl/structs.py:23-176) and of ``clean_pufferl.train`` (core.py:213-259) that sits
around ``c_gae.compute_gae``.

The reference keeps values / rewards / dones in host numpy arrays in *arrival order* (per step: the rows of the
non-masked envs, until ``batch_size`` rows are stored), sorts them by ``(env_id, step)`` with a Python ``sorted()`` over
131072 tuples, runs GAE on the host and copies the advantages back.  Here everything stays in HBM in a fixed ``[T, N]``
layout (+ mask); the reference's arrival indices and its sorted order are reproduced with two prefix sums, and the scan
runs in ``phc_gae`` -- no host round trip, one sync per rollout (the size of the ragged result).

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

from .c_gae import compute_gae_cuda


class RolloutBuffer:
    def __init__(self, num_envs: int, batch_size: int, max_steps: Optional[int] = None, obs_dim: int = 0, device="cuda"):
        self.N, self.batch_size = int(num_envs), int(batch_size)
        # masked rows make a rollout longer than batch_size / num_envs steps; leave room (grown on demand)
        self.T = int(max_steps) if max_steps is not None else 2 * (-(-self.batch_size // self.N)) + 2
        self.device = torch.device(device)
        f = dict(dtype=torch.float32, device=self.device)
        self.values = torch.zeros(self.T, self.N, **f)
        self.rewards = torch.zeros(self.T, self.N, **f)
        self.dones = torch.zeros(self.T, self.N, **f)
        self.truncateds = torch.zeros(self.T, self.N, **f)
        self.mask = torch.zeros(self.T, self.N, dtype=torch.bool, device=self.device)
        self.obs = torch.zeros(self.T, self.N, obs_dim, **f) if obs_dim else None
        self.step = 0
        self._stored = torch.zeros((), dtype=torch.int64, device=self.device)    # rows stored so far (device counter, no sync)
        self.idxs = None

    # ---- structs.py:108-131 -------------------------------------------------------------------------------
    def store(self, value, reward, done, trunc, mask, obs=None) -> None:
        """One env step for all N envs (env_id = arange(N)); everything stays on the device, nothing syncs."""
        t = self.step
        if t >= self.T:
            self._grow()
        self.values[t].copy_(value)
        self.rewards[t].copy_(reward)
        self.dones[t].copy_(done)
        self.truncateds[t].copy_(trunc)
        self.mask[t].copy_(mask)
        if self.obs is not None and obs is not None:
            self.obs[t].copy_(obs)
        self._stored += mask.sum()
        self.step += 1

    def _grow(self) -> None:
        for name in ("values", "rewards", "dones", "truncateds", "mask", "obs"):
            a = getattr(self, name)
            if a is not None:
                setattr(self, name, torch.cat([a, torch.zeros_like(a)], 0))
        self.T *= 2

    @property
    def full(self) -> bool:                       # structs.py:104-106 (one device->host read)
        return int(self._stored) >= self.batch_size

    # ---- structs.py:133-145 -------------------------------------------------------------------------------
    def sort_training_data(self) -> torch.Tensor:
        """Arrival indices (the reference's row numbers) in (env_id, step) order, as a device int64 tensor."""
        T = self.step
        m = self.mask[:T]
        rank = torch.cumsum(m.reshape(-1).to(torch.int64), 0).reshape(T, self.N)      # arrival rank (1-based), step-major
        keep = m & (rank <= self.batch_size)                                          # store() stops at batch_size rows
        keep_em = keep.t().reshape(-1)                                                # env-major order = sorted by (env, step)
        pos = torch.nonzero(keep_em).squeeze(-1)                                      # the one sync of the rollout
        arrival = (rank - 1).t().reshape(-1)
        self.idxs = arrival[pos]
        self._pos_em = pos                                                            # positions in the env-major [N*T] flattening
        return self.idxs

    def _sorted(self, a: torch.Tensor) -> torch.Tensor:
        return a[: self.step].t().reshape(-1)[self._pos_em]

    # ---- core.py:213-259 ----------------------------------------------------------------------------------
    def compute_advantages(self, gamma: float, gae_lambda: float, extra_reward: Optional[torch.Tensor] = None):
        """Advantages and returns in the reference's sorted order (length = number of stored rows)."""
        if self.idxs is None:
            self.sort_training_data()
        d, v, r = self._sorted(self.dones), self._sorted(self.values), self._sorted(self.rewards)
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
        self.mask.zero_()
        self.idxs = None
