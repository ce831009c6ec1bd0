"""TEST / BASELINE INFRASTRUCTURE -- a batched torch (CPU, fp32) restatement of the reference's hot path.

Purpose: the *timed CPU baseline* (bench.py ``cpu_baseline`` and ``--impl reference``).  The reference runs this
path as a long sequence of eager torch ops over ``[N, 24, k]`` tensors; this file restates the same op sequences
(gather -> lerp/slerp -> heading-frame rotations -> exp rewards -> termination test -> running-norm) so that its
cost on the host cores is representative of the reference's own implementation, which cannot travel to the GPU
box (it needs the reference checkout).  It is validated against the reference-generated golden vectors in
tests/test_torch_port.py.  The bit-exact parity oracle is oracle/phc_oracle.c, not this file.

Only tests/, __graft_entry__.smoke() and bench.py may import it.  Citations are to the reference checkout.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

Tensor = torch.Tensor


# ---- puffer_phc/torch_utils.py -------------------------------------------------------------------
def q_mul(a: Tensor, b: Tensor) -> Tensor:                     # torch_utils.py:55-75
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    ww = (az + ax) * (bx + by)
    yy = (aw - ay) * (bw + bz)
    zz = (aw + ay) * (bw - bz)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (az - ax) * (bx - by))
    return torch.stack((qq - xx + (ax + aw) * (bx + bw), qq - yy + (aw - ax) * (by + bz),
                        qq - zz + (az + ay) * (bw - bx), qq - ww + (az - ay) * (by - bz)), -1)


def q_conj(a: Tensor) -> Tensor:                                # torch_utils.py:79-82
    return torch.cat((-a[..., :3], a[..., 3:]), -1)


def q_rotate(q: Tensor, v: Tensor) -> Tensor:                   # torch_utils.py:274-281
    w = q[..., 3:]
    qv = q[..., :3]
    a = v * (2.0 * w * w - 1.0)
    b = torch.cross(qv, v, dim=-1) * w * 2.0
    c = qv * (qv * v).sum(-1, keepdim=True) * 2.0
    return a + b + c


def tan_norm(q: Tensor) -> Tensor:                              # torch_utils.py:285-297
    ex = torch.zeros_like(q[..., :3]); ex[..., 0] = 1
    ez = torch.zeros_like(q[..., :3]); ez[..., 2] = 1
    return torch.cat((q_rotate(q, ex), q_rotate(q, ez)), -1)


def heading(q: Tensor) -> Tensor:                               # torch_utils.py:369-380
    ex = torch.zeros_like(q[..., :3]); ex[..., 0] = 1
    d = q_rotate(q, ex)
    return torch.atan2(d[..., 1], d[..., 0])


def q_from_angle_z(angle: Tensor) -> Tensor:                    # torch_utils.py:354-358 with axis z (:384-408)
    th = (angle / 2).unsqueeze(-1)
    zeros = torch.zeros_like(th)
    q = torch.cat((zeros, zeros, th.sin(), th.cos()), -1)
    return q / q.norm(p=2, dim=-1, keepdim=True).clamp(min=1e-9)


def q_angle(q: Tensor) -> Tensor:                               # torch_utils.py:86-106 (angle only)
    w = q[..., 3]
    s = torch.sqrt(1 - w * w)
    ang = 2 * torch.acos(w)
    ang = torch.atan2(torch.sin(ang), torch.cos(ang))
    return torch.where(torch.abs(s) > 1e-5, ang, torch.zeros_like(ang))


def q_exp_map(q: Tensor) -> Tensor:                             # torch_utils.py:144-150
    w = q[..., 3]
    s = torch.sqrt(1 - w * w)
    ang = 2 * torch.acos(w)
    ang = torch.atan2(torch.sin(ang), torch.cos(ang))
    mask = torch.abs(s) > 1e-5
    axis = q[..., :3] / s.unsqueeze(-1)
    default = torch.zeros_like(axis); default[..., 2] = 1
    ang = torch.where(mask, ang, torch.zeros_like(ang))
    axis = torch.where(mask.unsqueeze(-1), axis, default)
    return ang.unsqueeze(-1) * axis


def slerp(q0: Tensor, q1: Tensor, t: Tensor) -> Tensor:         # torch_utils.py:110-131
    c = (q0 * q1).sum(-1)
    q1 = torch.where((c < 0).unsqueeze(-1), -q1, q1)
    c = c.abs().unsqueeze(-1)
    h = torch.acos(c)
    s = torch.sqrt(1.0 - c * c)
    out = torch.sin((1 - t) * h) / s * q0 + torch.sin(t * h) / s * q1
    out = torch.where(s.abs() < 0.001, 0.5 * q0 + 0.5 * q1, out)
    return torch.where(c.abs() >= 1, q0, out)


# ---- puffer_phc/motion_lib.py ----------------------------------------------------------------------
def frame_blend(time: Tensor, length: Tensor, nf: Tensor, dt: Tensor) -> Tuple[Tensor, Tensor, Tensor]:   # :655-665
    time = time.clone()
    phase = torch.clip(time / length, 0.0, 1.0)
    time[time < 0] = 0
    i0 = (phase * (nf - 1)).long()
    i1 = torch.min(i0 + 1, nf - 1)
    blend = torch.clip((time - i0 * dt) / dt, 0.0, 1.0)
    return i0, i1, blend


def motion_state(T: Dict[str, Tensor], ids: Tensor, times: Tensor, offset=None, full: bool = True) -> Dict[str, Tensor]:   # :549-626
    i0, i1, blend = frame_blend(times, T["motion_len"][ids], T["num_frames"][ids], T["motion_dt"][ids])
    f0, f1 = i0 + T["length_starts"][ids], i1 + T["length_starts"][ids]
    b = blend.view(-1, 1, 1)
    pos = (1.0 - b) * T["gts"][f0] + b * T["gts"][f1]
    if offset is not None:
        pos = pos + offset[:, None, :]
    out = {
        "rg_pos": pos,
        "rb_rot": slerp(T["grs"][f0], T["grs"][f1], b),
        "body_vel": (1.0 - b) * T["gvs"][f0] + b * T["gvs"][f1],
        "body_ang_vel": (1.0 - b) * T["gavs"][f0] + b * T["gavs"][f1],
    }
    if full:
        local = slerp(T["lrs"][f0], T["lrs"][f1], b)
        out["dof_pos"] = q_exp_map(local[:, 1:]).reshape(len(ids), -1)
        out["dof_vel"] = ((1.0 - b) * T["dvs"][f0] + b * T["dvs"][f1]).reshape(len(ids), -1)
        out["motion_aa"] = T["motion_aa"][f0]
        out["root_pos"], out["root_rot"] = pos[:, 0].clone(), out["rb_rot"][:, 0].clone()
        out["root_vel"], out["root_ang_vel"] = out["body_vel"][:, 0].clone(), out["body_ang_vel"][:, 0].clone()
        out["motion_bodies"], out["motion_limb_weights"] = T["motion_bodies"][ids], T["limb_weights"][ids]
    return out


# ---- puffer_phc/envs/common.py -----------------------------------------------------------------------
def self_obs(pos: Tensor, rot: Tensor, vel: Tensor, ang: Tensor) -> Tensor:      # :23-103 with the env's constant flags
    N, J = pos.shape[:2]
    hinv = q_from_angle_z(-heading(rot[:, 0])).unsqueeze(1).expand(N, J, 4).reshape(-1, 4)
    local = q_rotate(hinv, (pos - pos[:, :1]).reshape(-1, 3)).reshape(N, -1)[:, 3:]
    rot_obs = tan_norm(q_mul(hinv, rot.reshape(-1, 4))).reshape(N, -1)
    v = q_rotate(hinv, vel.reshape(-1, 3)).reshape(N, -1)
    a = q_rotate(hinv, ang.reshape(-1, 3)).reshape(N, -1)
    return torch.cat((pos[:, 0, 2:3], local, rot_obs, v, a), -1)


def task_obs(pos, rot, vel, ang, rpos, rrot, rvel, rang) -> Tensor:                  # :106-176, time_steps=1, upright
    N, J = pos.shape[:2]
    hd = heading(rot[:, 0])
    hinv = q_from_angle_z(-hd).unsqueeze(1).expand(N, J, 4).reshape(-1, 4)
    h = q_from_angle_z(hd).unsqueeze(1).expand(N, J, 4).reshape(-1, 4)
    d_pos = q_rotate(hinv, (rpos - pos).reshape(-1, 3))
    d_rot = q_mul(q_mul(hinv, q_mul(rrot, q_conj(rot)).reshape(-1, 4)), h)
    d_vel = q_rotate(hinv, (rvel - vel).reshape(-1, 3))
    d_ang = q_rotate(hinv, (rang - ang).reshape(-1, 3))
    l_pos = q_rotate(hinv, (rpos - pos[:, :1]).reshape(-1, 3))
    l_rot = tan_norm(q_mul(hinv, rrot.reshape(-1, 4)))
    parts = (d_pos, tan_norm(d_rot), d_vel, d_ang, l_pos, l_rot)
    return torch.cat([p.reshape(N, -1) for p in parts], -1)


def reward(pos, rot, vel, ang, rpos, rrot, rvel, rang, k, w) -> Tuple[Tensor, Tensor]:   # :270-322
    d_pos = ((rpos - pos) ** 2).mean(-1).mean(-1)
    d_rot = (q_angle(q_mul(rrot, q_conj(rot))) ** 2).mean(-1)
    d_vel = ((rvel - vel) ** 2).mean(-1).mean(-1)
    d_ang = ((rang - ang) ** 2).mean(-1).mean(-1)
    r = [torch.exp(-k[i] * d) for i, d in enumerate((d_pos, d_rot, d_vel, d_ang))]
    return w[0] * r[0] + w[1] * r[1] + w[2] * r[2] + w[3] * r[3], torch.stack(r, -1)


def im_reset(progress, pos, rpos, pass_time, term_dist, use_mean=False) -> Tuple[Tensor, Tensor]:    # :325-364
    dist = torch.norm(pos - rpos, dim=-1)
    if use_mean:
        fallen = torch.any(dist.mean(-1, keepdim=True) > term_dist[0], dim=-1)
    else:
        fallen = torch.any(dist > term_dist, dim=-1)
    fallen = fallen & (progress > 1)
    return torch.where(pass_time, torch.ones_like(fallen), fallen), fallen


# ---- the post-physics step (puffer_phc/envs/humanoid_phc.py:136-149) -------------------------------------
def step(T, S, dt=1.0 / 30.0, k=(100.0, 10.0, 0.1, 0.1), w=(0.5, 0.3, 0.1, 0.1), power_coef=0.0005, term=0.25) -> Dict[str, Tensor]:
    st = S["body_state"][:, :24]
    pos, rot, vel, ang = st[..., 0:3], st[..., 3:7], st[..., 7:10], st[..., 10:13]
    ids, prog = S["motion_ids"], S["progress"]
    t0 = prog * dt + S["start_time"] + S["start_offset"]                              # :1233-1235
    t1 = (prog + 1) * dt + S["start_time"] + S["start_offset"]                        # :1060-1064
    r0 = motion_state(T, ids, t0, S["global_offset"])                                 # the reference computes the full dict
    r1 = motion_state(T, ids, t1, S["global_offset"])
    rew, raw = reward(pos, rot, vel, ang, r0["rg_pos"], r0["rb_rot"], r0["body_vel"], r0["body_ang_vel"], k, w)
    if "dof_force" in S:                                                              # :1295-1303
        pr = -power_coef * torch.abs(S["dof_force"] * S["dof_vel"]).sum(-1)
        pr[prog <= 3] = 0
        rew = rew + pr
        raw = torch.cat((raw, pr[:, None]), -1)
    pass_time = t0 >= T["motion_len"][ids]                                            # :1315
    reset, term_flag = im_reset(prog, pos.clone(), r0["rg_pos"].clone(), pass_time, torch.full((24,), term))
    obs = torch.cat((self_obs(pos, rot, vel, ang),
                     task_obs(pos, rot, vel, ang, r1["rg_pos"], r1["rb_rot"], r1["body_vel"], r1["body_ang_vel"])), -1)   # :947
    return {"obs": obs, "reward": rew, "reward_raw": raw, "reset": reset, "terminated": term_flag}


# ---- puffer_phc/policies/running_norm.py -------------------------------------------------------------------
def rms_forward(x, mean, var, eps=1e-5, clip=10.0):                                  # :15-20
    return torch.clamp((x - mean) / torch.sqrt(var + eps), -clip, clip)


def rms_update(x, mean, var, count):                                                 # :23-34
    wgt = 1 / count
    return mean * (1 - wgt) + x.mean(0, keepdim=True) * wgt, var * (1 - wgt) + x.var(0, unbiased=False, keepdim=True) * wgt, count + 1
