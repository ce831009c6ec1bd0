"""Drop-ins for the four hot free functions of the reference's ``puffer_phc/envs/common.py``: same names,
positional arguments, return shapes/dtypes.  Inputs may be strided views of the PhysX rigid-body buffer
(``state[..., :24, 0:3]``, reference puffer_phc/envs/humanoid_phc.py:546-549) -- they are consumed in place
through ``phc_view`` (base pointer + env/body strides); only tensors whose innermost stride is not 1 are
copied.  Kernels: csrc/imitation.cu.  CUDA tensors only; there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from .. import _ffi

_KEYS_K = ("k_pos", "k_rot", "k_vel", "k_ang_vel")
_KEYS_W = ("w_pos", "w_rot", "w_vel", "w_ang_vel")


def _prep(*ts):
    _ffi.require_cuda(*ts)
    return [_ffi.as_view_tensor(t) for t in ts]


def compute_imitation_observations_v6(root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel, ref_body_pos,
                                      ref_body_rot, ref_body_vel, ref_body_ang_vel, time_steps: int, upright: bool, out=None):
    """reference envs/common.py:106-176 -> ``[B, time_steps*J*24]`` (per future step six body-major blocks 3J|6J|3J|3J|3J|6J).  With
    ``time_steps > 1`` the four reference tensors hold ``B*time_steps*J`` bodies, viewed as ``[B, time_steps, J, .]`` like the reference
    does.  ``out`` (optional, beyond the reference's signature): a ``[B, >= time_steps*J*24]`` float32 buffer with unit inner stride to
    write into instead of allocating."""
    lib = _ffi.load()
    time_steps = int(time_steps)
    if time_steps < 1:
        raise ValueError("compute_imitation_observations_v6: time_steps must be >= 1")
    B, J = body_pos.shape[0], body_pos.shape[1]
    if time_steps > 1:                      # [B, time_steps, J, k] / [B*time_steps, J, k] / any shape with that many elements -> rows
        ref_body_pos, ref_body_vel, ref_body_ang_vel = (x.reshape(B * time_steps, J, 3) for x in (ref_body_pos, ref_body_vel, ref_body_ang_vel))
        ref_body_rot = ref_body_rot.reshape(B * time_steps, J, 4)
    ts = _prep(root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel, ref_body_pos, ref_body_rot, ref_body_vel,
               ref_body_ang_vel)
    obs = torch.empty((B, 24 * J * time_steps), dtype=torch.float32, device=ts[2].device) if out is None else out
    with _ffi.on_device(obs.device):
        _ffi.check(lib.phc_imitation_obs_v6(*[_ffi.view3(t) for t in ts], B, J, time_steps, int(bool(upright)),
                                            _ffi.ptr(obs), obs.stride(0), _ffi.stream_ptr()), "compute_imitation_observations_v6")
    return obs


def compute_humanoid_observations_smpl_max(body_pos, body_rot, body_vel, body_ang_vel, smpl_params, limb_weight_params,
                                           local_root_obs, root_height_obs, upright, has_smpl_params, has_limb_weight_params):
    """reference envs/common.py:23-103 -> ``[B, (1) + 3(J-1) + 12J (+ params)]``."""
    lib = _ffi.load()
    ts = _prep(body_pos, body_rot, body_vel, body_ang_vel)
    B, J = ts[0].shape[0], ts[0].shape[1]
    W = (1 if root_height_obs else 0) + 3 * (J - 1) + 12 * J
    extra = []
    if has_smpl_params:            # plain pass-through columns (common.py:96-100)
        extra.append(smpl_params)
    if has_limb_weight_params:
        extra.append(limb_weight_params)
    Wt = W + sum(int(x.shape[-1]) for x in extra)
    obs = torch.empty((B, Wt), dtype=torch.float32, device=ts[0].device)
    with _ffi.on_device(obs.device):
        _ffi.check(lib.phc_self_obs_smpl_max(*[_ffi.view3(t) for t in ts], B, J, int(bool(local_root_obs)),
                                             int(bool(root_height_obs)), int(bool(upright)), _ffi.ptr(obs), obs.stride(0),
                                             _ffi.stream_ptr()), "compute_humanoid_observations_smpl_max")
    col = W
    for x in extra:
        obs[:, col:col + x.shape[-1]] = x
        col += x.shape[-1]
    return obs


def dof_to_obs_smpl(pose):
    """reference envs/common.py:179-189: exp-map triplets -> tan-norm, ``[B, 3*nj] -> [B, 6*nj]`` (through the AMP kernel)."""
    B, jts = pose.shape
    z3, z4 = pose.new_zeros(B, 3), pose.new_zeros(B, 4)
    z4[:, 3] = 1
    full = build_amp_observations_smpl(z3, z4, z3, z3, _pad69(pose), _pad69(pose), pose.new_zeros(B, 0, 3), pose.new_zeros(B, 0),
                                       pose.new_zeros(B, 0), torch.arange(jts, device=pose.device), True, False, True, False, False, True)
    return full[:, 12:12 + 2 * jts].contiguous()


def _pad69(x):
    if x.shape[1] == 69:
        return x
    out = x.new_zeros(x.shape[0], 69)
    out[:, : x.shape[1]] = x
    return out


def build_amp_observations_smpl(root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos, shape_params,
                                limb_weight_params, dof_subset, local_root_obs, root_height_obs, has_dof_subset, has_shape_obs_disc,
                                has_limb_weight_obs, upright):
    """reference envs/common.py:192-267 -> ``[B, (1) + 6 + 3 + 3 + 6*nj + 3*nj + 3*K (+ shape) (+ limb)]``."""
    lib = _ffi.load()
    _ffi.require_cuda(root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos)
    ts = [t.to(torch.float32).contiguous() for t in (root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos)]
    B, K = ts[0].shape[0], ts[6].shape[1]
    if ts[4].shape[1] != 69 or ts[5].shape[1] != 69:
        raise ValueError("build_amp_observations_smpl: dof_pos / dof_vel must be [B, 69] (23 SMPL joints)")
    sub = None
    nj = 23
    if has_dof_subset:
        sub = dof_subset.to(device=ts[0].device, dtype=torch.int64).contiguous()
        if sub.numel() % 3:
            raise ValueError("dof_subset must list whole joints (a multiple of 3 indices)")
        nj = sub.numel() // 3
    W = (1 if root_height_obs else 0) + 12 + 9 * nj + 3 * K
    extra = []
    if has_shape_obs_disc:
        extra.append(shape_params)
    if has_limb_weight_obs:
        extra.append(limb_weight_params)
    Wt = W + sum(int(x.shape[-1]) for x in extra)
    obs = torch.empty((B, Wt), dtype=torch.float32, device=ts[0].device)
    with _ffi.on_device(obs.device):
        _ffi.check(lib.phc_amp_obs_smpl(*[_ffi.ptr(t) for t in ts], _ffi.ptr(sub), nj, K, int(bool(local_root_obs)),
                                        int(bool(root_height_obs)), int(bool(upright)), B, _ffi.ptr(obs), obs.stride(0),
                                        _ffi.ref_device(), _ffi.stream_ptr()), "build_amp_observations_smpl")
    col = W
    for x in extra:
        obs[:, col:col + x.shape[-1]] = x
        col += x.shape[-1]
    return obs


def amp_obs_history_step(amp_obs_buf, root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos, shape_params,
                         limb_weight_params, dof_subset, local_root_obs, root_height_obs, has_dof_subset, has_shape_obs_disc,
                         has_limb_weight_obs, upright):
    """``HumanoidPHC._update_hist_amp_obs()`` followed by ``_compute_amp_observations()`` (reference puffer_phc/envs/humanoid_phc.py:
    154-157, 1123-1174, 1339-1348) on ``amp_obs_buf [N, num_amp_obs_steps, num_amp_obs_per_step]`` IN PLACE, in one kernel: every
    env's rows move one step back and row 0 receives the current observation (same arguments as ``build_amp_observations_smpl``)."""
    lib = _ffi.load()
    _ffi.require_cuda(amp_obs_buf, root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos)
    if amp_obs_buf.dim() != 3 or not amp_obs_buf.is_contiguous() or amp_obs_buf.dtype != torch.float32:
        raise TypeError("amp_obs_buf must be a contiguous float32 [N, steps, width] tensor (it is updated in place)")
    ts = [t.to(torch.float32).contiguous() for t in (root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos)]
    B, S, Wt = amp_obs_buf.shape
    K = ts[6].shape[1]
    sub, nj = None, 23
    if has_dof_subset:
        sub = dof_subset.to(device=ts[0].device, dtype=torch.int64).contiguous()
        nj = sub.numel() // 3
    W = (1 if root_height_obs else 0) + 12 + 9 * nj + 3 * K
    extra = ([shape_params] if has_shape_obs_disc else []) + ([limb_weight_params] if has_limb_weight_obs else [])
    if Wt != W + sum(int(x.shape[-1]) for x in extra):
        raise ValueError(f"amp_obs_buf rows are {Wt} wide, the observation is {W + sum(int(x.shape[-1]) for x in extra)}")
    with _ffi.on_device(amp_obs_buf.device):
        _ffi.check(lib.phc_amp_obs_hist_step(*[_ffi.ptr(t) for t in ts], _ffi.ptr(sub), nj, K, int(bool(local_root_obs)),
                                             int(bool(root_height_obs)), int(bool(upright)), B, _ffi.ptr(amp_obs_buf), S, Wt,
                                             _ffi.ref_device(), _ffi.stream_ptr()), "amp_obs_history_step")
    col = W
    for x in extra:                                   # plain pass-through columns of the current row (common.py:253-266)
        amp_obs_buf[:, 0, col:col + x.shape[-1]] = x
        col += x.shape[-1]
    return amp_obs_buf


def compute_imitation_reward(root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel, ref_body_pos, ref_body_rot,
                             ref_body_vel, ref_body_ang_vel, rwd_specs: Dict[str, float], out=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference envs/common.py:270-322 -> ``(reward [B], reward_raw [B,4])``.  root_pos/root_rot are unused there too.
    ``out`` (optional): ``(reward, reward_raw)`` buffers to write into."""
    lib = _ffi.load()
    ts = _prep(body_pos, body_rot, body_vel, body_ang_vel, ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel)
    B, J = ts[0].shape[0], ts[0].shape[1]
    k = (C.c_float * 4)(*[float(rwd_specs[n]) for n in _KEYS_K])
    w = (C.c_float * 4)(*[float(rwd_specs[n]) for n in _KEYS_W])
    if out is None:
        reward = torch.empty(B, dtype=torch.float32, device=ts[0].device)
        raw = torch.empty((B, 4), dtype=torch.float32, device=ts[0].device)
    else:
        reward, raw = out
    with _ffi.on_device(reward.device):
        _ffi.check(lib.phc_imitation_reward(*[_ffi.view3(t) for t in ts], B, J, k, w, _ffi.ptr(reward), _ffi.ptr(raw), raw.stride(0),
                                            _ffi.stream_ptr()), "compute_imitation_reward")
    return reward, raw


def compute_humanoid_im_reset(reset_buf, progress_buf, contact_buf, contact_body_ids, rigid_body_pos, ref_body_pos, pass_time,
                              enable_early_termination, termination_distance, use_mean):
    """reference envs/common.py:325-364 -> ``(reset, terminated)`` with reset_buf's dtype.  contact_buf and
    contact_body_ids are unused by the reference as well."""
    lib = _ffi.load()
    pos, ref = _prep(rigid_body_pos, ref_body_pos)
    _ffi.require_cuda(progress_buf, pass_time, termination_distance)
    B, J = pos.shape[0], pos.shape[1]
    prog = progress_buf if (progress_buf.dtype == torch.int16 and progress_buf.is_contiguous()) else progress_buf.to(torch.int16).contiguous()
    pt = pass_time if (pass_time.dtype == torch.bool and pass_time.is_contiguous()) else pass_time.to(torch.bool).contiguous()
    td = termination_distance
    if td.dtype != torch.float32 or td.dim() != 1 or not td.is_contiguous():
        td = td.to(torch.float32).reshape(-1).contiguous()
    if td.numel() == 1 and J > 1:
        td = td.expand(J).contiguous()
    flags = torch.empty((2, B), dtype=torch.bool, device=pos.device)     # one allocation for both outputs
    reset, term = flags[0], flags[1]
    with _ffi.on_device(pos.device):
        _ffi.check(lib.phc_im_reset(_ffi.ptr(prog), _ffi.view3(pos), _ffi.view3(ref), _ffi.ptr(pt), int(bool(enable_early_termination)),
                                    _ffi.ptr(td), int(bool(use_mean)), B, J, _ffi.ptr(reset), _ffi.ptr(term), _ffi.ref_device(),
                                    _ffi.stream_ptr()),
                   "compute_humanoid_im_reset")
    if reset_buf.dtype != torch.bool:
        reset, term = reset.to(reset_buf.dtype), term.to(reset_buf.dtype)
    return reset, term


def compute_mpjpe(rigid_body_pos, ref_body_pos):
    """``(body_pos - rg_pos).norm(dim=-1).mean(dim=-1)`` -- the per-env evaluation error HumanoidPHC.step puts in
    ``extras["mpjpe"]`` when ``flag_im_eval`` is set (reference envs/humanoid_phc.py:159-163; EvalStats, scripts/train.py:139-166)."""
    lib = _ffi.load()
    pos, ref = _prep(rigid_body_pos, ref_body_pos)
    B, J = pos.shape[0], pos.shape[1]
    out = torch.empty(B, dtype=torch.float32, device=pos.device)
    with _ffi.on_device(pos.device):
        _ffi.check(lib.phc_mpjpe(_ffi.view3(pos), _ffi.view3(ref), B, J, _ffi.ptr(out), _ffi.ref_device(), _ffi.stream_ptr()),
                   "compute_mpjpe")
    return out
