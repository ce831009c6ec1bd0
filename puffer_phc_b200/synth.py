"""Seeded synthetic inputs for the hot path: an AMASS-shaped motion library and PhysX-shaped sim states.

Host-side data preparation only (plain torch, any device); none of this is on the timed path.  The
layouts produced are exactly the ones the reference's loader hands to the hot path:

* motion tables as in ``MotionLibBase.load_motions`` (reference puffer_phc/motion_lib.py:396-420):
  ``gts [F,24,3]``, ``grs/lrs [F,24,4]`` (xyzw), ``gvs/gavs [F,24,3]``, ``dvs [F,23,3]``,
  ``motion_aa [F,72]`` and the per-motion vectors, ``length_starts`` = exclusive cumsum (:416-419);
* the sim state as Isaac Gym's rigid-body tensor, AoS ``[N, bodies, 13]`` = pos3, rot4, vel3, angvel3
  (reference puffer_phc/envs/humanoid_phc.py:542-549).

Recipe follows SURVEY.md section 8(d): frame counts ~ lognormal(ln 250, 0.8) clipped to [10, 300], 30 fps
(a configurable fraction of clips at other rates so that blend != 0 is exercised), smooth random local
rotations, toes/hands at identity, quaternion signs not canonicalised, runs of identical frames in every
``freeze_every``-th clip (exercises the slerp fall-back branches), FK over the SMPL tree.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

NUM_BODIES = 24
# SMPL kinematic tree of the reference humanoid (reference puffer_phc/assets/smpl_humanoid.xml via
# SkeletonTree.from_mjcf; body order = puffer_phc/body_sets.py:11-36).
SMPL_PARENTS = (-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22)
# Approximate bone offsets in metres (synthetic skeleton; only the shape of the data matters here).
SMPL_OFFSETS = (
    (0.0, 0.0, 0.0), (-0.007, 0.070, -0.091), (-0.005, 0.034, -0.375), (-0.044, -0.014, -0.398),
    (0.119, 0.026, -0.056), (-0.004, -0.068, -0.091), (-0.009, -0.038, -0.383), (-0.042, 0.016, -0.398),
    (0.123, -0.025, -0.048), (-0.027, -0.003, 0.109), (0.001, 0.006, 0.135), (0.025, 0.002, 0.053),
    (-0.043, -0.003, 0.214), (0.051, 0.005, 0.065), (-0.034, 0.079, 0.122), (-0.009, 0.091, 0.031),
    (-0.028, 0.260, -0.013), (-0.001, 0.249, 0.009), (-0.015, 0.084, -0.008), (-0.039, -0.082, 0.119),
    (-0.009, -0.096, 0.033), (-0.021, -0.254, -0.013), (-0.006, -0.255, 0.008), (-0.010, -0.085, -0.006),
)
IDENTITY_JOINTS = (4, 8, 18, 23)  # L_Toe, R_Toe, L_Hand, R_Hand keep identity local rotation

TABLE_KEYS = ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "motion_aa", "motion_len", "motion_dt",
              "num_frames", "length_starts", "motion_bodies", "limb_weights")


# ---- small quaternion helpers (xyzw) for data generation only -------------------------------------
def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack((aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw,
                        aw * bw - ax * bx - ay * by - az * bz), -1)


def _qconj(q):
    return torch.cat((-q[..., :3], q[..., 3:]), -1)


def _qrot(q, v):
    qv, w = q[..., :3], q[..., 3:]
    t = 2.0 * torch.cross(qv, v, dim=-1)
    return v + w * t + torch.cross(qv, t, dim=-1)


def _exp_to_quat(e):
    ang = e.norm(dim=-1, keepdim=True)
    half = 0.5 * ang
    k = torch.where(ang > 1e-8, torch.sin(half) / ang.clamp_min(1e-8), 0.5 - ang * ang / 48.0)
    return torch.cat((e * k, torch.cos(half)), -1)


def _quat_to_rotvec(q):
    q = torch.where(q[..., 3:] < 0, -q, q)
    s = q[..., :3].norm(dim=-1, keepdim=True)
    ang = 2.0 * torch.atan2(s, q[..., 3:])
    return torch.where(s > 1e-8, q[..., :3] / s.clamp_min(1e-8) * ang, 2.0 * q[..., :3])


def make_motion_library(num_clips: int = 11313, seed: int = 0, device="cpu", min_frames: int = 10,
                        max_frames: int = 300, median_frames: float = 250.0, sigma: float = 0.8,
                        other_fps_fraction: float = 0.0, freeze_every: int = 50) -> Dict[str, torch.Tensor]:
    """Build the synthetic library; returns a dict with TABLE_KEYS (+ ``motion_fps``) on ``device``."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    M = int(num_clips)
    nf = torch.exp(math.log(median_frames) + sigma * torch.randn(M, generator=g)).round().clamp(min_frames, max_frames).long()
    fps = torch.full((M,), 30.0, dtype=torch.float64)
    if other_fps_fraction > 0:
        pick = torch.rand(M, generator=g) < other_fps_fraction
        alt = torch.tensor([20.0, 60.0, 120.0], dtype=torch.float64)[torch.randint(0, 3, (M,), generator=g)]
        fps = torch.where(pick, alt, fps)
    # per-motion scalars exactly as the loader forms them (python double, then fp32): motion_lib.py:375-401
    motion_len = ((1.0 / fps) * (nf - 1).double()).float()
    motion_dt = (1.0 / fps).float()
    starts = torch.cumsum(nf, 0) - nf
    F = int(nf.sum())

    # per-clip smooth-rotation parameters: 3 sinusoids per joint axis
    amp = (torch.rand(M, NUM_BODIES, 3, 3, generator=g) * 2 - 1) * (0.6 / math.sqrt(3.0))
    frq = torch.rand(M, NUM_BODIES, 3, 3, generator=g) * 1.5
    phs = torch.rand(M, NUM_BODIES, 3, 3, generator=g) * (2 * math.pi)
    sign = torch.where(torch.rand(M, NUM_BODIES, generator=g) < 0.6, -1.0, 1.0)  # un-canonicalised signs
    yaw0 = (torch.rand(M, generator=g) * 2 - 1) * math.pi
    yaw_rate = (torch.rand(M, generator=g) * 2 - 1) * 0.8
    vel_amp = (torch.rand(M, 2, 2, generator=g) * 2 - 1) * 1.0
    vel_frq = torch.rand(M, 2, 2, generator=g) * 0.5 + 0.05
    vel_phs = torch.rand(M, 2, 2, generator=g) * (2 * math.pi)
    base_h = 0.85 + 0.1 * torch.rand(M, generator=g)
    bodies = torch.cat((torch.randint(0, 3, (M, 1), generator=g).float(), torch.randn(M, 16, generator=g)), 1)
    limb_w = torch.rand(M, 10, generator=g)
    # frozen runs: clips m % freeze_every == 0 hold one pose for frames [a, a+r)
    fr_a = (torch.rand(M, generator=g) * (nf - 1).float()).long()
    fr_r = torch.randint(2, 6, (M,), generator=g)
    frozen = (torch.arange(M) % max(freeze_every, 1) == 0) if freeze_every > 0 else torch.zeros(M, dtype=torch.bool)

    dev = torch.device(device)
    to = lambda x: x.to(dev)  # noqa: E731
    nf_d, starts_d, fps_d = to(nf), to(starts), to(fps.float())
    clip = torch.repeat_interleave(torch.arange(M, device=dev), nf_d)                  # [F]
    fidx = torch.arange(F, device=dev) - starts_d[clip]                                  # frame within clip
    a, r, fz = to(fr_a)[clip], to(fr_r)[clip], to(frozen)[clip]
    fidx_eff = torch.where(fz & (fidx >= a) & (fidx < a + r), a, fidx)
    t = fidx_eff.float() / fps_d[clip]                                                   # [F]

    out_lrs = torch.empty(F, NUM_BODIES, 4, device=dev)
    chunk = 262144
    amp_d, frq_d, phs_d, sign_d = to(amp), to(frq), to(phs), to(sign)
    yaw0_d, yawr_d = to(yaw0), to(yaw_rate)
    for s in range(0, F, chunk):
        c, tt = clip[s:s + chunk], t[s:s + chunk]
        e = (amp_d[c] * torch.sin(2 * math.pi * frq_d[c] * tt[:, None, None, None] + phs_d[c])).sum(-1)   # [n,24,3]
        e[:, 0] *= 0.25                                                                   # small root tilt
        q = _exp_to_quat(e)
        yaw = yaw0_d[c] + yawr_d[c] * tt
        qy = torch.stack((torch.zeros_like(yaw), torch.zeros_like(yaw), torch.sin(0.5 * yaw), torch.cos(0.5 * yaw)), -1)
        q[:, 0] = _qmul(qy, q[:, 0])
        q = q * sign_d[c][..., None]
        q[:, list(IDENTITY_JOINTS)] = torch.tensor([0.0, 0.0, 0.0, 1.0], device=dev)
        out_lrs[s:s + chunk] = q
    lrs = out_lrs

    # root translation: integral of a sum of velocity sinusoids (about 1 m/s), height wobble
    va, vf, vp = to(vel_amp)[clip], to(vel_frq)[clip], to(vel_phs)[clip]
    w = 2 * math.pi * vf
    xy = (va / w * (torch.cos(vp) - torch.cos(w * t[:, None, None] + vp))).sum(-1)       # [F,2]
    z = to(base_h)[clip] + 0.03 * torch.sin(2.0 * t + to(yaw0)[clip])
    root = torch.cat((xy, z[:, None]), -1)

    # forward kinematics over the SMPL tree
    offs = torch.tensor(SMPL_OFFSETS, device=dev)
    grs = torch.empty_like(lrs)
    gts = torch.empty(F, NUM_BODIES, 3, device=dev)
    grs[:, 0], gts[:, 0] = lrs[:, 0], root
    for j in range(1, NUM_BODIES):
        p = SMPL_PARENTS[j]
        grs[:, j] = _qmul(grs[:, p], lrs[:, j])
        gts[:, j] = gts[:, p] + _qrot(grs[:, p], offs[j].expand(F, 3))

    # finite-difference velocities inside each clip (one-sided at clip ends)
    first, last = fidx == 0, fidx == nf_d[clip] - 1
    ip = torch.where(last, torch.arange(F, device=dev), torch.arange(F, device=dev) + 1)
    im = torch.where(first, torch.arange(F, device=dev), torch.arange(F, device=dev) - 1)
    span = (ip - im).clamp_min(1).float() / fps_d[clip]
    gvs = (gts[ip] - gts[im]) / span[:, None, None]
    gavs = _quat_to_rotvec(_qmul(grs[ip], _qconj(grs[im]))) / span[:, None, None]
    # dof velocities from consecutive local rotations, last frame repeats (motion_lib.py:119-140)
    nxt = torch.where(last, torch.arange(F, device=dev), torch.arange(F, device=dev) + 1)
    dv = _quat_to_rotvec(_qmul(_qconj(lrs), lrs[nxt])) * fps_d[clip][:, None, None]
    prev = torch.where(last & ~first, torch.arange(F, device=dev) - 1, torch.arange(F, device=dev))
    dvs = dv[prev][:, 1:].contiguous()
    motion_aa = _quat_to_rotvec(lrs).reshape(F, 72).contiguous()

    return {
        "gts": gts.contiguous(), "grs": grs.contiguous(), "lrs": lrs.contiguous(), "gvs": gvs.contiguous(),
        "gavs": gavs.contiguous(), "dvs": dvs, "motion_aa": motion_aa,
        "motion_len": to(motion_len), "motion_dt": to(motion_dt), "motion_fps": fps_d,
        "num_frames": nf_d, "length_starts": starts_d, "motion_bodies": to(bodies), "limb_weights": to(limb_w),
    }


def make_env_state(tables: Dict[str, torch.Tensor], num_envs: int, seed: int = 1, bodies_per_env: int = NUM_BODIES,
                   motion_ids: Optional[torch.Tensor] = None, with_dof: bool = True) -> Dict[str, torch.Tensor]:
    """Per-env scalars and a PhysX-shaped rigid-body state near the reference pose.

    Returns ``motion_ids [N] i64``, ``progress [N] i16``, ``start_time/start_offset [N] f32``,
    ``global_offset [N,3]``, ``body_state [N, bodies_per_env, 13]`` and (optionally) ``dof_force/dof_vel [N,69]``.
    """
    dev = tables["gts"].device
    g = torch.Generator(device="cpu").manual_seed(seed)
    N, M = int(num_envs), int(tables["motion_len"].shape[0])
    ids = torch.randint(0, M, (N,), generator=g) if motion_ids is None else motion_ids.cpu().long()
    nf = tables["num_frames"].cpu()[ids]
    mdt = tables["motion_dt"].cpu()[ids]
    sim_dt = 1.0 / 30.0
    # start on a 1/30 s grid like sample_time_interval (motion_lib.py:526-535)
    mlen = tables["motion_len"].cpu()[ids]
    start = ((torch.rand(N, generator=g) * mlen) / sim_dt).long().float() * sim_dt
    remaining = ((mlen - start) / sim_dt).clamp_min(0).long()
    progress = (torch.rand(N, generator=g) * (remaining + 3).float()).long()
    early = torch.rand(N, generator=g) < 0.02
    progress = torch.where(early, torch.randint(0, 2, (N,), generator=g), progress).to(torch.int16)
    offset = torch.cat((torch.rand(N, 2, generator=g) * 2 - 1, torch.zeros(N, 1)), 1)
    # nearest table frame at the reward time t = progress*dt + start
    tnow = progress.float() * sim_dt + start
    fr = torch.minimum((tnow / mdt).round().long(), nf - 1).clamp_min(0) + tables["length_starts"].cpu()[ids]
    fr = fr.to(dev)
    wide = (torch.rand(N, generator=g) < 0.10).float()[:, None, None]
    sig_p = 0.05 * (1 - wide) + 0.2 * wide
    pos = tables["gts"][fr] + offset.to(dev)[:, None, :] + (torch.randn(N, NUM_BODIES, 3, generator=g) * sig_p).to(dev)
    rot = _qmul(tables["grs"][fr], _exp_to_quat((torch.randn(N, NUM_BODIES, 3, generator=g) * 0.1).to(dev)))
    exact = (torch.rand(N, generator=g) < 0.03).to(dev)                      # a few envs track the frame exactly
    rot = torch.where(exact[:, None, None], tables["grs"][fr], rot)
    vel = tables["gvs"][fr] + (torch.randn(N, NUM_BODIES, 3, generator=g) * 0.3).to(dev)
    ang = tables["gavs"][fr] + (torch.randn(N, NUM_BODIES, 3, generator=g) * 0.5).to(dev)
    state = torch.zeros(N, bodies_per_env, 13, device=dev)
    state[:, :NUM_BODIES] = torch.cat((pos, rot, vel, ang), -1)
    out = {
        "motion_ids": ids.to(dev), "progress": progress.to(dev), "start_time": start.to(dev),
        "start_offset": torch.zeros(N, device=dev), "global_offset": offset.to(dev), "body_state": state.contiguous(),
    }
    if with_dof:
        out["dof_force"] = (torch.randn(N, 69, generator=g) * 10.0).to(dev)
        out["dof_vel"] = (torch.randn(N, 69, generator=g) * 1.0).to(dev)
    return out


def make_rollout(num_envs: int = 4096, horizon: int = 32, seed: int = 2, device="cpu", p_done: float = 0.01):
    """Flat env-major rollout arrays for the GAE pass."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = num_envs * horizon
    values = torch.randn(L, generator=g)
    rewards = torch.rand(L, generator=g)
    dones = (torch.rand(L, generator=g) < p_done).float()
    return {"dones": dones.to(device), "values": values.to(device), "rewards": rewards.to(device)}
