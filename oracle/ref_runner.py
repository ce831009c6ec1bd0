"""TEST / BASELINE INFRASTRUCTURE -- runs the reference's OWN Python functions for the hot path.

The files are the unmodified reference sources that ``__graft_entry__.build()`` stages under ``oracle/_ref/`` (git-ignored,
shipped to the GPU box by gpurun; see oracle/build_ref.py).  They are imported from there and executed with torch eager /
TorchScript on whatever device the caller picks: ``cpu`` (the device the golden vectors were made on) or ``cuda`` (the device the
reference really runs this path on, reference puffer_phc/envs/humanoid_phc.py:875-897, 979, 1099, 1257, 1322).

Only tests/, __graft_entry__.smoke() and bench.py (baseline legs) may import this module.  ``HumanoidPHC`` itself needs Isaac Gym, so the
glue between the functions (motion times, pass_time, power reward, obs concatenation) is restated from humanoid_phc.py at the lines
cited in ``step``; everything with arithmetic in it is the reference's code.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings
from types import SimpleNamespace
from typing import Dict, Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

EVAL_BODY_IDS = [j for j in range(24) if j not in (4, 8, 18, 23)]     # body_sets.py:42,57
RWD = dict(k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1, w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1)  # config.py:25-32
POWER_COEF = 0.0005                                                     # config.py:96
DT = 1.0 / 30.0                                                         # isaacgym_env.py:39-41

_booted: Optional[SimpleNamespace] = None


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF, "puffer_phc", f)) for f in ("motion_lib.py", "torch_utils.py", "envs/common.py"))


def boot() -> SimpleNamespace:
    """Import the staged reference modules (only the un-vendored ``smpl_sim`` import is stubbed; the query
    path never touches it)."""
    global _booted
    if _booted is not None:
        return _booted
    if not available():
        raise RuntimeError("oracle/_ref does not hold the reference files; run __graft_entry__.build() where /root/reference exists")
    warnings.filterwarnings("ignore")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("smpl_sim", "smpl_sim.smpllib", "smpl_sim.smpllib.smpl_parser"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["smpl_sim.smpllib.smpl_parser"].SMPL_Parser = type("SMPL_Parser", (), {"__init__": lambda s, *a, **k: None})
    from puffer_phc import motion_lib as ml
    from puffer_phc import torch_utils
    from puffer_phc.envs import common
    from puffer_phc.poselib_skeleton import SkeletonTree
    spec = importlib.util.spec_from_file_location("ref_running_norm", os.path.join(REF, "puffer_phc", "policies", "running_norm.py"))
    rn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rn)
    c_gae = None
    try:
        import c_gae                                           # the reference's own .pyx, compiled into oracle/_ref/
    except Exception:
        pass
    _booted = SimpleNamespace(ml=ml, common=common, torch_utils=torch_utils, SkeletonTree=SkeletonTree, rn=rn, c_gae=c_gae, root=REF)
    return _booted


def lib_from_tables(T: Dict[str, torch.Tensor], device="cpu"):
    """A reference ``MotionLibSMPL`` whose tables are the given tensors (the query code only reads attributes, SURVEY.md
    section 8c), on ``device``."""
    R = boot()
    lib = object.__new__(R.ml.MotionLibSMPL)
    dev = torch.device(device)
    lib._device = dev
    lib._sim_fps = 1 / DT
    g = lambda k: T[k].to(dev)   # noqa: E731
    for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs"):
        setattr(lib, k, g(k))
    lib._motion_aa = g("motion_aa")
    lib._motion_lengths, lib._motion_dt = g("motion_len"), g("motion_dt")
    lib._motion_fps = g("motion_fps") if "motion_fps" in T else torch.round(1.0 / g("motion_dt"))
    lib._motion_num_frames, lib.length_starts = g("num_frames"), g("length_starts")
    lib._motion_bodies, lib._motion_limb_weights = g("motion_bodies"), g("limb_weights")
    return lib


def load_cmu(device="cpu"):
    """BASELINE config 1: the sample clip through the reference's own loader (host work), tables moved to ``device``."""
    R = boot()
    cfg = SimpleNamespace(motion_file=os.path.join(REF, "sample_data", "cmu_mocap_05_06.pkl"), device="cpu",
                          fix_height=R.ml.FixHeightMode.no_fix, min_length=5, max_length=300, im_eval=False,
                          num_thread=1, smpl_type="smpl", step_dt=DT, is_deterministic=True)
    lib = R.ml.MotionLibSMPL(cfg)
    lib.mesh_parsers = None
    sk = R.SkeletonTree.from_mjcf(os.path.join(REF, "puffer_phc", "assets", "smpl_humanoid.xml"))
    import numpy as np
    lib.load_motions(skeleton_trees=[sk], gender_betas=torch.zeros(1, 17), limb_weights=np.zeros((1, 10)), random_sample=False)
    T = {"gts": lib.gts, "grs": lib.grs, "lrs": lib.lrs, "gvs": lib.gvs, "gavs": lib.gavs, "dvs": lib.dvs,
         "motion_aa": lib._motion_aa, "motion_len": lib._motion_lengths, "motion_dt": lib._motion_dt, "motion_fps": lib._motion_fps,
         "num_frames": lib._motion_num_frames, "length_starts": lib.length_starts, "motion_bodies": lib._motion_bodies,
         "limb_weights": lib._motion_limb_weights}
    return lib_from_tables(T, device), T


def step(lib, S: Dict[str, torch.Tensor], eval_mode: bool = False, power: bool = True, with_blend: bool = False,
         full_state: bool = False, eval_distance: float = 0.5) -> Dict[str, torch.Tensor]:
    """The post-physics half of ``HumanoidPHC.step`` (humanoid_phc.py:136-149) on the reference's functions, on the device the
    inputs live on.  ``S``: body_state [N,B,13], progress i16, start_time, start_offset, motion_ids, global_offset (+ dof_force /
    dof_vel).  Returns obs [N,934], reward, reward_raw, reset, terminated (+ optional extras)."""
    R = boot()
    common = R.common
    st = S["body_state"][:, :24]
    body_pos, body_rot, body_vel, body_ang = st[..., 0:3], st[..., 3:7], st[..., 7:10], st[..., 10:13]   # :546-549 (strided views)
    ids, prog = S["motion_ids"], S["progress"]
    dev = st.device
    out = {}
    t0 = prog * DT + S["start_time"] + S["start_offset"]                     # _compute_reward :1233-1235
    t1 = (prog + 1) * DT + S["start_time"] + S["start_offset"]               # _compute_task_obs :1060-1064
    r0 = lib.get_motion_state(ids, t0, offset=S["global_offset"])            # :1236 via _get_state_from_motionlib_cache :895
    r1 = lib.get_motion_state(ids, t1, offset=S["global_offset"])            # :1066
    if with_blend:
        for tag, tt in (("t0", t0), ("t1", t1)):
            i0, i1, bl = lib._calc_frame_blend(tt, lib._motion_lengths[ids], lib._motion_num_frames[ids], lib._motion_dt[ids])
            out[f"{tag}_idx0"], out[f"{tag}_idx1"], out[f"{tag}_blend"] = i0, i1, bl
    if full_state:
        for tag, r in (("t0", r0), ("t1", r1)):
            for k, v in r.items():
                out[f"{tag}_{k}"] = v
    # reward :1257-1270, power :1295-1303
    rew, raw = common.compute_imitation_reward(body_pos[:, 0], body_rot[:, 0], body_pos, body_rot, body_vel, body_ang,
                                               r0["rg_pos"], r0["rb_rot"], r0["body_vel"], r0["body_ang_vel"], RWD)
    if power:
        pw = torch.abs(torch.multiply(S["dof_force"], S["dof_vel"])).sum(dim=-1)
        power_reward = -POWER_COEF * pw
        power_reward[prog <= 3] = 0
        out["reward"] = rew + power_reward
        out["reward_raw"] = torch.cat([raw, power_reward[:, None]], -1)
    else:
        out["reward"], out["reward_raw"] = rew, raw
    # reset :1311-1333 (train: all 24 bodies, 0.25 m, any; eval :1424-1435: 20 bodies, 0.5 m, mean)
    pass_time = t0 >= lib._motion_lengths[ids]
    reset_buf = torch.ones(len(ids), dtype=torch.bool, device=dev)
    contact = torch.zeros(len(ids), 24, 3, device=dev)
    cids = torch.zeros(4, dtype=torch.long, device=dev)
    if eval_mode:
        td = torch.full((24,), float(eval_distance), device=dev)
        rs, tm = common.compute_humanoid_im_reset(reset_buf, prog, contact, cids, body_pos[..., EVAL_BODY_IDS, :],
                                                  r0["rg_pos"][..., EVAL_BODY_IDS, :], pass_time, True, td[..., EVAL_BODY_IDS], True)
    else:
        td = torch.full((24,), 0.25, device=dev)
        rs, tm = common.compute_humanoid_im_reset(reset_buf, prog, contact, cids, body_pos, r0["rg_pos"], pass_time, True, td, False)
    out["reset"], out["terminated"], out["pass_time"] = rs, tm, pass_time
    # observations :947, :979-991, :1099-1112
    self_obs = common.compute_humanoid_observations_smpl_max(body_pos, body_rot, body_vel, body_ang, None, None,
                                                             True, True, True, False, False)
    task_obs = common.compute_imitation_observations_v6(body_pos[:, 0], body_rot[:, 0], body_pos, body_rot, body_vel, body_ang,
                                                        r1["rg_pos"], r1["rb_rot"], r1["body_vel"], r1["body_ang_vel"], 1, True)
    out["obs"] = torch.cat([self_obs, task_obs], dim=-1)
    return out


def flag_margins(lib, S: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Distance of every flag-deciding quantity from its threshold, computed with the reference's expressions on the inputs'
    device: per env ``min_j | ||body_j - ref_j|| - 0.25 |`` (common.py:347-350) and ``|t - motion_len|`` (humanoid_phc.py:1315)."""
    st = S["body_state"][:, :24]
    ids, prog = S["motion_ids"], S["progress"]
    t0 = prog * DT + S["start_time"] + S["start_offset"]
    r0 = lib.get_motion_state(ids, t0, offset=S["global_offset"])
    d = torch.norm(st[..., 0:3] - r0["rg_pos"], dim=-1)
    return {"dist": d, "dist_margin": (d - 0.25).abs().min(dim=-1).values, "time_margin": (t0 - lib._motion_lengths[ids]).abs()}


def rollout_tail(obs: torch.Tensor, mean: torch.Tensor, var: torch.Tensor):
    """RunningNorm.forward (policies/running_norm.py:15-20) with the given statistics, through the reference's module."""
    R = boot()
    rn = R.rn.RunningNorm(obs.shape[1]).to(obs.device)
    rn.running_mean.copy_(mean.reshape(1, -1))
    rn.running_var.copy_(var.reshape(1, -1))
    return rn, rn(obs)
