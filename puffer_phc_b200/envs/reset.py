"""Reset path ("next" row f1 of SURVEY.md section 8): what runs every step on the envs whose ``reset_buf`` is set.

Reference flow: ``PHCPufferEnv.step`` (reference puffer_phc/clean_pufferl/env.py:114-140) -> ``HumanoidPHC.reset`` ->
``_reset_envs`` (puffer_phc/envs/humanoid_phc.py:663-674) -> ``_reset_ref_state_init`` (:692-727: ``_sample_ref_state``
:843-873 then ``_set_env_state`` :899-929) -> ``_reset_env_tensors`` (:729-777) -> ``_compute_observations(env_ids)``.

Here ``_sample_ref_state`` + ``_set_env_state`` are ONE kernel (``phc_reset_ref_state``: the motion-state query of each
reset env is written straight into the root / dof / rigid-body state tensors instead of being materialised as 10 tensors
and scattered with 12 indexed copies); the observation of the reset envs uses the stand-alone observation kernels on the
gathered rows, like the reference.  Random numbers stay on torch's generator (same call, same shape, same device as the
reference: ``torch.rand(len(env_ids))`` inside ``sample_time_interval``), so sampled start times are identical.
Isaac Gym's ``set_*_tensor_indexed`` / ``refresh_*`` calls stay the reference's (they are simulator API, not arithmetic).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _ffi
from . import common


@dataclass
class EnvTensors:
    """The per-env tensors ``HumanoidPHC`` owns (humanoid_phc.py:523-598), same names without the underscore."""
    rigid_body_state: torch.Tensor            # [N, bodies_per_env, 13]  PhysX AoS
    humanoid_root_states: torch.Tensor        # [N, 13]
    dof_pos: torch.Tensor                     # [N, 69]
    dof_vel: torch.Tensor                     # [N, 69]
    progress_buf: torch.Tensor                # [N] int16
    reset_buf: torch.Tensor                   # [N] bool
    terminate_buf: torch.Tensor               # [N] bool
    global_offset: torch.Tensor               # [N, 3]
    motion_start_times: torch.Tensor          # [N]
    motion_start_times_offset: torch.Tensor   # [N]
    sampled_motion_ids: torch.Tensor          # [N] int64
    obs_buf: torch.Tensor                     # [N, 934]


def reset_envs(env: EnvTensors, motion_lib, env_ids: torch.Tensor, random_start: bool = True, flag_test: bool = False,
               dt: float = 1.0 / 30.0, motion_times: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``HumanoidPHC._reset_envs(env_ids)`` for ``StateInit.Random`` (``random_start``) / ``StateInit.Start``.
    Returns the sampled motion start times of the reset envs.  ``motion_times`` overrides the sampling (tests)."""
    if env_ids.numel() == 0:
        return env_ids.new_zeros(0, dtype=torch.float32)
    lib = _ffi.load()
    _ffi.require_cuda(env_ids, env.rigid_body_state)
    env_ids = env_ids.to(torch.int64).contiguous()
    ids = env.sampled_motion_ids[env_ids]
    # _sample_ref_state (:843-857)
    if motion_times is None:
        if random_start:
            motion_times = motion_lib.sample_time_interval(ids)                       # _sample_time (:838-841)
        else:
            motion_times = torch.zeros(env_ids.shape[0], device=env_ids.device)
        if flag_test:
            motion_times[:] = 0
    motion_times = motion_times.to(torch.float32).contiguous()
    bs = env.rigid_body_state
    assert bs.is_contiguous() and env.humanoid_root_states.is_contiguous() and env.dof_pos.is_contiguous() and env.dof_vel.is_contiguous()
    # _sample_ref_state query (offset = the envs' CURRENT global offset, :859-861) + _set_env_state (:899-929)
    with torch.cuda.device(bs.device):
        _ffi.check(lib.phc_reset_ref_state(C.byref(motion_lib.ctables), _ffi.ptr(env_ids), _ffi.ptr(env.sampled_motion_ids),
                                           _ffi.ptr(motion_times), _ffi.ptr(env.global_offset), env_ids.shape[0],
                                           _ffi.ptr(env.humanoid_root_states), _ffi.ptr(env.dof_pos), _ffi.ptr(env.dof_vel), _ffi.ptr(bs),
                                           bs.stride(0), _ffi.ref_device(), _ffi.stream_ptr()), "phc_reset_ref_state")
    # _reset_ref_state_init tail (:721-727) and _reset_env_tensors (:774-777)
    env.global_offset[env_ids] = 0
    env.motion_start_times[env_ids] = motion_times
    env.motion_start_times_offset[env_ids] = 0
    env.progress_buf[env_ids] = 0
    env.reset_buf[env_ids] = 0
    env.terminate_buf[env_ids] = 0
    # _compute_observations(env_ids) (:935-959; _compute_humanoid_obs :961-991, _compute_task_obs :1048-1112)
    st = bs[env_ids][:, :24]
    bp, br, bv, ba = st[..., 0:3], st[..., 3:7], st[..., 7:10], st[..., 10:13]
    t1 = (env.progress_buf[env_ids] + 1) * dt + env.motion_start_times[env_ids] + env.motion_start_times_offset[env_ids]
    ref = motion_lib.get_motion_state(ids, t1, env.global_offset[env_ids], keys=("rg_pos", "rb_rot", "body_vel", "body_ang_vel"))
    self_obs = common.compute_humanoid_observations_smpl_max(bp, br, bv, ba, None, None, True, True, True, False, False)
    task_obs = common.compute_imitation_observations_v6(bp[:, 0], br[:, 0], bp, br, bv, ba, ref["rg_pos"], ref["rb_rot"], ref["body_vel"],
                                                        ref["body_ang_vel"], 1, True)
    env.obs_buf[env_ids] = torch.cat([self_obs, task_obs], dim=-1)
    return motion_times


def auto_reset(env: EnvTensors, motion_lib, terminals: torch.Tensor, truncations: torch.Tensor, masks: torch.Tensor,
               episode_returns: torch.Tensor, episode_lengths: torch.Tensor, rewards: torch.Tensor, **kw):
    """The device part of ``PHCPufferEnv.step`` after ``env.step`` (clean_pufferl/env.py:111-140): reset the flagged envs and
    derive terminals / truncations / masks and the episode statistics.  Returns (reset_indices, finished returns, lengths)."""
    terminals[:] = False
    truncations[:] = False
    masks[:] = True
    reset_flags = env.reset_buf.clone()
    terminated = env.terminate_buf.clone()                        # extras["terminate"] (humanoid_phc.py:151)
    reset_indices = torch.nonzero(reset_flags).squeeze(-1)
    fin_ret = episode_returns[reset_indices].clone()
    fin_len = episode_lengths[reset_indices].clone()
    if reset_indices.numel() > 0:
        reset_envs(env, motion_lib, reset_indices, **kw)
        episode_returns[reset_indices] = 0
        episode_lengths[reset_indices] = 0
        terminals[:] = terminated                                 # env.py:124-126
        trunc = reset_flags & ~terminated                         # env.py:128-129
        truncations[:] = trunc
        masks[:] = ~trunc                                         # env.py:132-133
    # env.py:139-140 reads reset_buf AFTER the reset cleared it, so every env accumulates
    episode_returns[~env.reset_buf] += rewards[~env.reset_buf]
    episode_lengths[~env.reset_buf] += 1
    return reset_indices, fin_ret, fin_len
