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
    with _ffi.on_device(bs.device):
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


class AutoReset:
    """The same auto-reset as ``auto_reset`` / ``reset_envs`` above, but on the device and without a host round trip
    (``phc_auto_reset``, csrc/reset_tail.cu): no ``torch.nonzero``, no indexed assignments, two kernel launches that can be captured
    in one CUDA graph together with the step.

    Semantics kept from the reference (puffer_phc/clean_pufferl/env.py:102-140, puffer_phc/envs/humanoid_phc.py:663-777, 843-959):
    the k-th flagged env in ascending order consumes the k-th uniform of ``phase`` (the reference draws ``torch.rand(len(env_ids))``;
    here ``torch.rand(N)`` is drawn up front because the count is not known on the host), the state query uses the env's current
    global offset which is then cleared, the observation row of a reset env is recomputed against the reference at ``t + 1`` of the
    new episode, terminals / truncations / masks and the episode returns / lengths follow env.py line by line, and the logged
    quantities are kept as device-resident sums (``metrics``; one read when the host wants to log).

    With ``fused`` (a ``FusedStep`` that accumulates RunningNorm moments with ``defer_moments=True``) the moments are corrected to
    the rows the reference's ``Experience`` really stores: the post-reset row of a terminated env, no row of a truncated env.
    """

    def __init__(self, env: EnvTensors, motion_lib, dt: float = 1.0 / 30.0, random_start: bool = True, flag_test: bool = False,
                 ref_device=None, obs_norm: Optional[torch.Tensor] = None, rms=None, fused=None, step_metrics: bool = True):
        self.lib = _ffi.load()
        self.env, self.motion_lib = env, motion_lib
        N = env.progress_buf.shape[0]
        dev = env.rigid_body_state.device
        self.N, self.device = N, dev
        _ffi.require_cuda(env.rigid_body_state, env.obs_buf, env.progress_buf)
        for name in ("rigid_body_state", "humanoid_root_states", "dof_pos", "dof_vel", "progress_buf", "reset_buf", "terminate_buf",
                     "global_offset", "motion_start_times", "motion_start_times_offset", "sampled_motion_ids", "obs_buf"):
            t = getattr(env, name)
            if t is not None and not t.is_contiguous():
                raise ValueError(f"AutoReset: env.{name} must be contiguous (it is updated in place)")
        if env.progress_buf.dtype != torch.int16 or env.sampled_motion_ids.dtype != torch.int64:
            raise TypeError("AutoReset: progress_buf must be int16 and sampled_motion_ids int64 (as the reference env holds them)")
        if env.reset_buf.dtype != torch.bool or env.terminate_buf.dtype != torch.bool:
            raise TypeError("AutoReset: reset_buf / terminate_buf must be torch.bool")
        self.terminals = torch.zeros(N, dtype=torch.bool, device=dev)          # env.py:41-43
        self.truncations = torch.zeros(N, dtype=torch.bool, device=dev)
        self.masks = torch.ones(N, dtype=torch.bool, device=dev)
        self.episode_returns = torch.zeros(N, dtype=torch.float32, device=dev)   # env.py:52-53
        self.episode_lengths = torch.zeros(N, dtype=torch.int32, device=dev)
        self.reset_ids = torch.empty(N, dtype=torch.int64, device=dev)
        self.reset_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.fused = fused
        if fused is not None and fused.accumulate_moments and not fused.defer_moments:
            raise ValueError("AutoReset: moment correction needs FusedStep(defer_moments=True)")
        if fused is not None and fused.accumulate_moments:
            fused.tail_slots_used = True
        self.metrics = fused.stats[1 + 2 * 934:] if fused is not None else torch.zeros(_ffi.NUM_METRICS, dtype=torch.float64, device=dev)
        self.step_metrics = bool(step_metrics) and not (fused is not None and fused.metrics)
        self._scratch = torch.empty(int(self.lib.phc_auto_reset_scratch_bytes(N)) // 8 + 1, dtype=torch.float64, device=dev)
        self.obs_norm, self.rms = obs_norm, rms
        if obs_norm is not None and rms is None:
            raise ValueError("AutoReset: obs_norm needs the RunningNorm")
        bs = env.rigid_body_state.reshape(N, -1)
        self._cenv = _ffi.ResetEnv(
            bs.data_ptr(), bs.stride(0), _ffi.ptr(env.humanoid_root_states).value, _ffi.ptr(env.dof_pos).value, _ffi.ptr(env.dof_vel).value,
            env.progress_buf.data_ptr(), env.motion_start_times.data_ptr(), env.motion_start_times_offset.data_ptr(),
            env.global_offset.data_ptr(), env.sampled_motion_ids.data_ptr(), env.reset_buf.data_ptr(), env.terminate_buf.data_ptr(),
            env.obs_buf.data_ptr(), env.obs_buf.stride(0), _ffi.ptr(obs_norm).value,
            rms.running_mean.data_ptr() if obs_norm is not None else None, rms.running_var.data_ptr() if obs_norm is not None else None)
        self._ccfg = _ffi.ResetCfg(float(torch.tensor(dt, dtype=torch.float32)), 0 if random_start else 1, int(bool(flag_test)),
                                   _ffi.ref_device(ref_device), float(rms.epsilon) if rms is not None else 1e-5,
                                   float(rms.clip) if rms is not None else 10.0)

    def __call__(self, rewards: torch.Tensor, reward_raw: Optional[torch.Tensor] = None, phase: Optional[torch.Tensor] = None):
        """Run the auto-reset for the flags currently in ``env.reset_buf`` / ``env.terminate_buf``.  ``rewards`` = this step's
        ``rew_buf``; ``phase`` = ``torch.rand(N)`` on the device (drawn here when omitted -- the RNG stays torch's).
        Returns ``(terminals, truncations, masks)`` (device bool tensors owned by this object)."""
        N = self.N
        _ffi.require_cuda(rewards, reward_raw, phase)
        if phase is None:
            phase = torch.rand(N, device=self.device)
        if phase.dtype != torch.float32 or not phase.is_contiguous() or phase.numel() < N:
            raise TypeError("AutoReset: phase must be a contiguous float32 tensor with N elements")
        if rewards.dtype != torch.float32 or not rewards.is_contiguous():
            raise TypeError("AutoReset: rewards must be a contiguous float32 [N] tensor")
        raw_dim = 0
        if reward_raw is not None:
            if reward_raw.dtype != torch.float32 or reward_raw.stride(1) != 1:
                raise TypeError("AutoReset: reward_raw must be float32 with unit inner stride")
            raw_dim = reward_raw.shape[1]
        book = _ffi.ResetBook(rewards.data_ptr(), _ffi.ptr(reward_raw).value, reward_raw.stride(0) if reward_raw is not None else 0, raw_dim,
                              self.terminals.data_ptr(), self.truncations.data_ptr(), self.masks.data_ptr(),
                              self.episode_returns.data_ptr(), self.episode_lengths.data_ptr(), self.metrics.data_ptr(),
                              1 if self.step_metrics else 0)
        f = self.fused
        mom = f is not None and f.accumulate_moments
        with _ffi.on_device(self.device):
            _ffi.check(self.lib.phc_auto_reset(C.byref(self.motion_lib.ctables), C.byref(self._cenv), C.byref(book), C.byref(self._ccfg),
                                               _ffi.ptr(phase), N, _ffi.ptr(self.reset_ids), _ffi.ptr(self.reset_count), _ffi.ptr(self._scratch),
                                               f.partials[f.num_partials:].data_ptr() if mom else None,
                                               f.row_adjust.data_ptr() if mom else None, _ffi.stream_ptr()), "phc_auto_reset")
        return self.terminals, self.truncations, self.masks

    def metric_values(self, reset: bool = False):
        """Device-resident sums as a dict (one device->host read): steps, reward, r_*, resets, terminations, truncations,
        episode_return / episode_length / episodes (means: divide by steps resp. episodes)."""
        v = self.metrics.cpu().tolist()
        if reset:
            self.metrics.zero_()
        return {k: v[i] for i, k in enumerate(_ffi.METRIC_NAMES)}
