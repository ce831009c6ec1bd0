"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE ITSELF.

Run in the build container only (the reference checkout is read from /root/reference and does not
exist on the GPU box):

    python tests/golden/make_golden.py [--ref /root/reference]

The reference has no tests or golden vectors of its own, so parity is pinned by
running its own torch (CPU, fp32) / Cython functions on seeded inputs and committing inputs + outputs:

  cmu_tables.npz    motion tables the reference loader builds from sample_data/cmu_mocap_05_06.pkl
  cmu_step.npz      BASELINE config 1 at N=256: per-env inputs and the reference outputs of
                    _calc_frame_blend, get_motion_state (t and t+1), compute_humanoid_observations_smpl_max,
                    compute_imitation_observations_v6, compute_imitation_reward (+ power term),
                    compute_humanoid_im_reset (train and eval variants)
  synth_tables.npz  a tiny synthetic library (mixed fps, frozen runs) from puffer_phc_b200.synth
  synth_step.npz    the same outputs on it
  gae.npz           c_gae.compute_gae on 512x32 and edge cases (the reference .pyx compiled by oracle/Makefile)
  rms.npz           RunningNorm.update x2 + forward
  sample_time.npz   sample_time_interval / get_motion_num_steps arithmetic
  amp.npz           build_amp_observations_smpl (AMP discriminator observation, off by default in the reference)
  mpjpe.npz         extras["mpjpe"] of the evaluation step on both step fixtures
  loader.npz        raw clips (the sample clip + three synthetic ones) and the tables load_motions builds from them

The glue between the functions (motion_times, pass_time, obs concatenation, power reward) is restated
from puffer_phc/envs/humanoid_phc.py at the lines cited below, because HumanoidPHC itself needs Isaac Gym.
"""
import argparse
import importlib.util
import os
import sys
import types
import warnings
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

EVAL_BODY_IDS = [j for j in range(24) if j not in (4, 8, 18, 23)]     # body_sets.py:42,57
K = dict(k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1, w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1)  # config.py:25-32
POWER_COEF = 0.0005                                                     # config.py:96
DT = 1.0 / 30.0                                                         # isaacgym_env.py:39-41 (sim 60 Hz x 2 substeps)


def boot_reference(ref):
    """stub the un-vendored smpl_sim import, then import the reference modules."""
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ref)
    for name in ("smpl_sim", "smpl_sim.smpllib", "smpl_sim.smpllib.smpl_parser"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["smpl_sim.smpllib.smpl_parser"].SMPL_Parser = type("SMPL_Parser", (), {"__init__": lambda s, *a, **k: None})
    from puffer_phc import motion_lib as ml
    from puffer_phc.envs import common
    from puffer_phc.poselib_skeleton import SkeletonTree
    spec = importlib.util.spec_from_file_location("ref_running_norm", f"{ref}/puffer_phc/policies/running_norm.py")
    rn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rn)
    return ml, common, SkeletonTree, rn


def load_cmu(ml, SkeletonTree, ref):
    cfg = SimpleNamespace(motion_file=f"{ref}/sample_data/cmu_mocap_05_06.pkl", device="cpu",
                          fix_height=ml.FixHeightMode.no_fix, min_length=5, max_length=300, im_eval=False,
                          num_thread=1, smpl_type="smpl", step_dt=DT, is_deterministic=True)
    lib = ml.MotionLibSMPL(cfg)
    lib.mesh_parsers = None
    sk = SkeletonTree.from_mjcf(f"{ref}/puffer_phc/assets/smpl_humanoid.xml")
    lib.load_motions(skeleton_trees=[sk], gender_betas=torch.zeros(1, 17), limb_weights=np.zeros((1, 10)),
                     random_sample=False)
    return lib


def lib_from_tables(ml, T):
    """Bypass load_motions: the query code only reads attributes."""
    lib = object.__new__(ml.MotionLibSMPL)
    lib._device = "cpu"
    lib._sim_fps = 1 / DT
    for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs"):
        setattr(lib, k, T[k])
    lib._motion_aa = T["motion_aa"]
    lib._motion_lengths, lib._motion_dt = T["motion_len"], T["motion_dt"]
    lib._motion_fps = T["motion_fps"]
    lib._motion_num_frames, lib.length_starts = T["num_frames"], T["length_starts"]
    lib._motion_bodies, lib._motion_limb_weights = T["motion_bodies"], T["limb_weights"]
    return lib


def tables_of(lib):
    return {
        "gts": lib.gts, "grs": lib.grs, "lrs": lib.lrs, "gvs": lib.gvs, "gavs": lib.gavs, "dvs": lib.dvs,
        "motion_aa": lib._motion_aa, "motion_len": lib._motion_lengths, "motion_dt": lib._motion_dt,
        "motion_fps": lib._motion_fps, "num_frames": lib._motion_num_frames, "length_starts": lib.length_starts,
        "motion_bodies": lib._motion_bodies, "limb_weights": lib._motion_limb_weights,
    }


def reference_step(lib, common, S):
    """The post-physics half of HumanoidPHC.step (humanoid_phc.py:136-149) on the reference functions."""
    st = S["body_state"][:, :24]
    body_pos, body_rot, body_vel, body_ang = st[..., 0:3], st[..., 3:7], st[..., 7:10], st[..., 10:13]  # :546-549
    ids, prog = S["motion_ids"], S["progress"]
    out = {}
    # _compute_reward :1233-1235
    t0 = prog * DT + S["start_time"] + S["start_offset"]
    # _compute_task_obs :1060-1064
    t1 = (prog + 1) * DT + S["start_time"] + S["start_offset"]
    out["t0"], out["t1"] = t0, t1
    for tag, tt in (("t0", t0), ("t1", t1)):
        i0, i1, bl = lib._calc_frame_blend(tt, lib._motion_lengths[ids], lib._motion_num_frames[ids], lib._motion_dt[ids])
        out[f"{tag}_idx0"], out[f"{tag}_idx1"], out[f"{tag}_blend"] = i0, i1, bl
        res = lib.get_motion_state(ids, tt, offset=S["global_offset"])
        for k, v in res.items():
            out[f"{tag}_{k}"] = v
    res_nooff = lib.get_motion_state(ids, t0, offset=None)
    out["t0_rg_pos_nooffset"] = res_nooff["rg_pos"]
    out["t0_root_pos_smpl"] = lib.get_root_pos_smpl(ids, t0)["root_pos"]
    r0 = {k: out[f"t0_{k}"] for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel")}
    r1 = {k: out[f"t1_{k}"] for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel")}
    # reward :1257-1270, power :1295-1303
    rew, raw = common.compute_imitation_reward(body_pos[:, 0], body_rot[:, 0], body_pos, body_rot, body_vel, body_ang,
                                               r0["rg_pos"], r0["rb_rot"], r0["body_vel"], r0["body_ang_vel"], K)
    out["reward_nopower"], out["reward_raw4"] = rew.clone(), raw.clone()
    power = torch.abs(torch.multiply(S["dof_force"], S["dof_vel"])).sum(dim=-1)
    power_reward = -POWER_COEF * power
    power_reward[prog <= 3] = 0
    out["reward"] = rew + power_reward
    out["reward_raw"] = torch.cat([raw, power_reward[:, None]], -1)
    # reset :1311-1333 (train: all 24 bodies, 0.25 m, max; eval :1424-1435: 20 bodies, 0.5 m, mean)
    pass_time = t0 >= lib._motion_lengths[ids]
    out["pass_time"] = pass_time
    reset_buf = torch.ones(len(ids), dtype=torch.bool)
    contact = torch.zeros(len(ids), 24, 3)
    for tag, bids, dist, use_mean in (("train", list(range(24)), 0.25, False), ("eval", EVAL_BODY_IDS, 0.5, True)):
        td = torch.full((24,), dist)
        rs, tm = common.compute_humanoid_im_reset(reset_buf, prog, contact, torch.zeros(4, dtype=torch.long),
                                                  body_pos[..., bids, :].clone(), r0["rg_pos"][..., bids, :].clone(),
                                                  pass_time, True, td[..., bids], use_mean)
        out[f"reset_{tag}"], out[f"terminated_{tag}"] = rs, tm
    rs, tm = common.compute_humanoid_im_reset(reset_buf, prog, contact, torch.zeros(4, dtype=torch.long), body_pos.clone(),
                                              r0["rg_pos"].clone(), pass_time, False, torch.full((24,), 0.25), False)
    out["reset_noearly"], out["terminated_noearly"] = rs, tm
    # observations :947, :979-991, :1099-1112
    self_obs = common.compute_humanoid_observations_smpl_max(body_pos, body_rot, body_vel, body_ang, None, None,
                                                             True, True, True, False, False)
    task_obs = common.compute_imitation_observations_v6(body_pos[:, 0], body_rot[:, 0], body_pos, body_rot, body_vel, body_ang,
                                                        r1["rg_pos"], r1["rb_rot"], r1["body_vel"], r1["body_ang_vel"], 1, True)
    out["obs"] = torch.cat([self_obs, task_obs], dim=-1)
    # non-default flag variants of the two observation functions (small slices)
    n = 32
    out["self_obs_variant"] = common.compute_humanoid_observations_smpl_max(
        body_pos[:n], body_rot[:n], body_vel[:n], body_ang[:n], None, None, False, False, False, False, False)
    out["task_obs_notupright"] = common.compute_imitation_observations_v6(
        body_pos[:n, 0], body_rot[:n, 0], body_pos[:n], body_rot[:n], body_vel[:n], body_ang[:n],
        r1["rg_pos"][:n], r1["rb_rot"][:n], r1["body_vel"][:n], r1["body_ang_vel"][:n], 1, False)
    return out


def save(name, d):
    arrs = {}
    for k, v in d.items():
        a = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        arrs[k] = a
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(arrs)} arrays")


def make_amp(common):
    """AMP observations (row f3): build_amp_observations_smpl (envs/common.py:192-267) on the inputs of cmu_step.npz.
    Reads the committed fixture, so it can be regenerated on its own (--only amp)."""
    z = np.load(os.path.join(HERE, "cmu_step.npz"))
    st = torch.from_numpy(z["in_body_state"])[:, :24]
    root_pos, root_rot, root_vel, root_ang = st[:, 0, 0:3], st[:, 0, 3:7], st[:, 0, 7:10], st[:, 0, 10:13]
    dof_pos, dof_vel = torch.from_numpy(z["t0_dof_pos"]), torch.from_numpy(z["t0_dof_vel"])
    dof_pos = dof_pos + 0.05 * torch.randn(dof_pos.shape, generator=torch.Generator().manual_seed(21))
    dof_pos[:8, 0:3] = 0.0                               # exercise the small-angle branch of exp_map_to_angle_axis
    dof_pos[8:16, 3:6] = torch.tensor([0.0, 0.0, 4.0])   # angle > pi: normalize_angle wraps
    key_ids = [7, 3, 21, 17]                             # R_Ankle, L_Ankle, R_Wrist, L_Wrist (body_sets.py:45)
    key_pos = st[:, key_ids, 0:3]
    joints = [j for j in range(23) if j not in (3, 7, 17, 22)]      # dof joints without toes / hands (humanoid_phc.py:186-196)
    subset = torch.tensor([3 * j + k for j in joints for k in range(3)], dtype=torch.long)
    shape = torch.randn(st.shape[0], 11, generator=torch.Generator().manual_seed(22))
    limb = torch.rand(st.shape[0], 10, generator=torch.Generator().manual_seed(23))
    out = {"root_pos": root_pos, "root_rot": root_rot, "root_vel": root_vel, "root_ang_vel": root_ang, "dof_pos": dof_pos,
           "dof_vel": dof_vel, "key_pos": key_pos, "subset": subset, "shape": shape, "limb": limb}
    variants = {  # local_root_obs, root_height_obs, has_dof_subset, has_shape_obs_disc, has_limb_weight_obs, upright
        "default": (True, True, True, False, False, True),
        "full_dof": (True, True, False, False, False, True),
        "global_root": (False, False, True, False, False, True),
        "not_upright": (True, True, True, False, False, False),
        "with_params": (True, True, True, True, True, True),
    }
    for name, f in variants.items():
        out[f"amp_{name}"] = common.build_amp_observations_smpl(root_pos, root_rot, root_vel, root_ang, dof_pos, dof_vel, key_pos,
                                                                 shape, limb, subset, *f)
        out[f"flags_{name}"] = np.array(f, dtype=np.int32)
    save("amp.npz", out)


def make_mpjpe():
    """Evaluation error of HumanoidPHC.step (humanoid_phc.py:159-163): the reference's torch expression on the committed
    sim states / reference positions of both step fixtures (--only mpjpe)."""
    out = {}
    for name in ("cmu_step", "synth_step"):
        z = np.load(os.path.join(HERE, name + ".npz"))
        body_pos = torch.from_numpy(z["in_body_state"])[:, :24, 0:3]
        rg_pos = torch.from_numpy(z["t0_rg_pos"])
        out[name] = (body_pos - rg_pos).norm(dim=-1).mean(dim=-1)
    save("mpjpe.npz", out)


def make_loader(ml, SkeletonTree, ref):
    """Motion table build (row f4): raw clips in the on-disk format of scripts/convert_amass_data.py:186-196 and the tables
    the reference loader (load_motions / load_motion_with_skeleton, motion_lib.py:257-429, 744-825) builds from them:
    the real sample clip plus three synthetic clips (one longer than max_length, one at 60 fps)."""
    import joblib
    from puffer_phc_b200 import synth
    sk = SkeletonTree.from_mjcf(f"{ref}/puffer_phc/assets/smpl_humanoid.xml")
    raw = joblib.load(f"{ref}/sample_data/cmu_mocap_05_06.pkl")
    clips = [raw[k] for k in raw]
    T = synth.make_motion_library(3, seed=31, min_frames=20, max_frames=60, median_frames=45.0, other_fps_fraction=0.0, freeze_every=2)
    for m in range(3):
        a, n = int(T["length_starts"][m]), int(T["num_frames"][m])
        clips.append({"root_trans_offset": T["gts"][a:a + n, 0].double().clone(), "pose_aa": T["motion_aa"][a:a + n].double().numpy(),
                      "pose_quat_global": T["grs"][a:a + n].double().numpy(), "beta": np.zeros(16), "gender": "neutral",
                      "fps": 60 if m == 2 else 30})
    max_length = 50
    cfg = SimpleNamespace(motion_file="", device="cpu", fix_height=ml.FixHeightMode.no_fix, min_length=5, max_length=max_length,
                          im_eval=False, num_thread=1, smpl_type="smpl", step_dt=DT, is_deterministic=True)
    lib = object.__new__(ml.MotionLibSMPL)
    lib.m_cfg, lib._device, lib._sim_fps, lib.mesh_parsers = cfg, "cpu", 1 / DT, None
    lib._motion_data_list = np.array(clips, dtype=object)
    lib._motion_data_keys = np.array([f"clip{i}" for i in range(len(clips))])
    lib._num_unique_motions = len(clips)
    lib.setup_constants(fix_height=ml.FixHeightMode.no_fix, num_thread=1)
    M = len(clips)
    lib.load_motions(skeleton_trees=[sk] * M, gender_betas=torch.zeros(M, 17), limb_weights=np.zeros((M, 10)), random_sample=False)
    out = {f"tab_{k}": v for k, v in tables_of(lib).items()}
    out["max_length"] = np.array(max_length)
    out["parents"] = sk.parent_indices.numpy()
    out["local_translation"] = sk.local_translation.numpy()
    for i, c in enumerate(clips):
        out[f"clip{i}_root_trans_offset"] = np.asarray(c["root_trans_offset"], dtype=np.float64)
        out[f"clip{i}_pose_aa"] = np.asarray(c["pose_aa"], dtype=np.float64)
        out[f"clip{i}_pose_quat_global"] = np.asarray(c["pose_quat_global"], dtype=np.float64)
        out[f"clip{i}_fps"] = np.array(c["fps"])
    # ---- the non-deterministic path: random crop start (random.randint) and random heading (np.random.random, scipy
    # Rotation) per clip, motion_lib.py:773-799.  The reference rotates pose_aa IN PLACE through a numpy view, so it runs
    # on deep copies; the seeds are part of the fixture.
    import copy
    import random
    clips_rnd = copy.deepcopy(clips)
    cfg_rnd = SimpleNamespace(**{**vars(cfg), "is_deterministic": False})
    lib_rnd = object.__new__(ml.MotionLibSMPL)
    lib_rnd.m_cfg, lib_rnd._device, lib_rnd._sim_fps, lib_rnd.mesh_parsers = cfg_rnd, "cpu", 1 / DT, None
    lib_rnd._motion_data_list = np.array(clips_rnd, dtype=object)
    lib_rnd._motion_data_keys = np.array([f"clip{i}" for i in range(len(clips))])
    lib_rnd._num_unique_motions = len(clips)
    lib_rnd.setup_constants(fix_height=ml.FixHeightMode.no_fix, num_thread=1)
    random.seed(7)
    np.random.seed(7)
    lib_rnd.load_motions(skeleton_trees=[sk] * M, gender_betas=torch.zeros(M, 17), limb_weights=np.zeros((M, 10)),
                         random_sample=False, sample_idxes=torch.arange(M))
    out.update({f"rnd_{k}": v for k, v in tables_of(lib_rnd).items()})
    out["rnd_seed"] = np.array(7)
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
    path = os.path.join(HERE, "loader.npz")
    np.savez_compressed(path, **arrs)                      # keeps float64 inputs (save() would cast them to float32)
    print(f"wrote loader.npz: {os.path.getsize(path) / 1e6:.2f} MB, {len(arrs)} arrays")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default=None, help="regenerate a single fixture (amp | loader | mpjpe)")
    args = ap.parse_args()
    torch.set_num_threads(1)
    ml, common, SkeletonTree, rn = boot_reference(args.ref)
    from puffer_phc_b200 import synth
    if args.only == "amp":
        make_amp(common)
        return
    if args.only == "mpjpe":
        make_mpjpe()
        return
    if args.only == "loader":
        make_loader(ml, SkeletonTree, args.ref)
        return

    # ---- config 1: the real clip -------------------------------------------------------------
    lib = load_cmu(ml, SkeletonTree, args.ref)
    T = tables_of(lib)
    save("cmu_tables.npz", T)
    S = synth.make_env_state(T, 256, seed=1, bodies_per_env=24)
    cmu = reference_step(lib, common, S)
    save("cmu_step.npz", {**{f"in_{k}": v for k, v in S.items()}, **cmu})

    # ---- tiny synthetic library: mixed fps, frozen runs, more bodies per env than 24 ------------
    T2 = synth.make_motion_library(12, seed=7, min_frames=10, max_frames=40, median_frames=25.0,
                                   other_fps_fraction=0.5, freeze_every=3)
    lib2 = lib_from_tables(ml, T2)
    save("synth_tables.npz", T2)
    S2 = synth.make_env_state(T2, 256, seed=11, bodies_per_env=26)
    S2["start_offset"] = (torch.rand(256, generator=torch.Generator().manual_seed(5)) - 0.3) * 0.05   # off-grid times
    S2["progress"][:4] = torch.tensor([0, 1, 2, 3], dtype=torch.int16)
    S2["start_time"][:2] = -0.5                                                                  # negative time branch
    syn = reference_step(lib2, common, S2)
    save("synth_step.npz", {**{f"in_{k}": v for k, v in S2.items()}, **syn})

    # ---- GAE: the reference's own compiled c_gae --------------------------------------------
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import c_gae
    gae = {}
    R = synth.make_rollout(512, 32, seed=2)
    gae["a_dones"], gae["a_values"], gae["a_rewards"] = (R[k].numpy() for k in ("dones", "values", "rewards"))
    gae["a_adv"] = c_gae.compute_gae(gae["a_dones"], gae["a_values"], gae["a_rewards"], 0.98, 0.2)
    gae["a_gamma_lambda"] = np.array([0.98, 0.2], np.float32)
    R = synth.make_rollout(1, 1000, seed=3, p_done=0.03)
    gae["b_dones"], gae["b_values"], gae["b_rewards"] = (R[k].numpy() for k in ("dones", "values", "rewards"))
    gae["b_adv"] = c_gae.compute_gae(gae["b_dones"], gae["b_values"], gae["b_rewards"], 0.99, 0.95)
    gae["b_gamma_lambda"] = np.array([0.99, 0.95], np.float32)
    for tag, L in (("c", 1), ("d", 2), ("e", 33)):
        R = synth.make_rollout(1, L, seed=4 + L, p_done=0.2)
        gae[f"{tag}_dones"], gae[f"{tag}_values"], gae[f"{tag}_rewards"] = (R[k].numpy() for k in ("dones", "values", "rewards"))
        gae[f"{tag}_adv"] = c_gae.compute_gae(gae[f"{tag}_dones"], gae[f"{tag}_values"], gae[f"{tag}_rewards"], 0.98, 0.2)
        gae[f"{tag}_gamma_lambda"] = np.array([0.98, 0.2], np.float32)
    save("gae.npz", gae)

    # ---- RunningNorm -------------------------------------------------------------------------
    norm = rn.RunningNorm(934)
    x1, x2 = cmu["obs"], syn["obs"][:200]
    norm.update(x1)
    rms = {"mean1": norm.running_mean.clone(), "var1": norm.running_var.clone(), "count1": norm.count.clone()}
    norm.update(x2)
    rms.update({"mean2": norm.running_mean.clone(), "var2": norm.running_var.clone(), "count2": norm.count.clone()})
    rms["fwd_rows"] = np.arange(0, 256, 16)
    big = x1[rms["fwd_rows"]].clone()
    big[0, :10] = 1e6
    big[1, :10] = -1e6                                   # exercise the +-clip
    rms["fwd_in"] = big
    rms["fwd_out"] = norm.forward(big)
    fresh = rn.RunningNorm(934)
    rms["fwd_out_fresh"] = fresh.forward(big)             # mean 0, var 1 initial state
    save("rms.npz", rms)

    # ---- sampling arithmetic -------------------------------------------------------------------
    lib3 = lib_from_tables(ml, T2)
    ids = torch.randint(0, 12, (4096,), generator=torch.Generator().manual_seed(9))
    torch.manual_seed(123)
    tt = lib3.sample_time_interval(ids)
    torch.manual_seed(123)
    phase = torch.rand(ids.shape)
    samp = {"ids": ids, "phase": phase, "time_interval": tt, "num_steps_all": lib3.get_motion_num_steps(),
            "motion_length_ids": lib3.get_motion_length(ids[:64])}
    # (get_motion_num_steps(ids) is broken in the reference itself -- motion_lib.py:547 divides the gathered
    #  frame counts by the un-gathered fps vector -- so only the ids=None form is pinned.)
    torch.manual_seed(321)
    samp["time_interval_trunc"] = lib_from_tables(ml, {k: v.clone() for k, v in T2.items()}).sample_time_interval(ids, truncate_time=0.1)
    torch.manual_seed(321)
    samp["phase_trunc"] = torch.rand(ids.shape)
    save("sample_time.npz", samp)
    make_amp(common)
    make_loader(ml, SkeletonTree, args.ref)
    make_mpjpe()


if __name__ == "__main__":
    main()
