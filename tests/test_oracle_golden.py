"""Pin the C oracle (oracle/phc_oracle.c) against fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only.  Integers and flags: bit-exact.  Floats: 1e-5 relative."""
import numpy as np
import pytest

from conftest import assert_close, assert_equal
from oracle import c_oracle as co

K = [100.0, 10.0, 0.1, 0.1]
W = [0.5, 0.3, 0.1, 0.1]
EVAL_IDS = [j for j in range(24) if j not in (4, 8, 18, 23)]
EVAL_MASK = sum(1 << j for j in EVAL_IDS)
CASES = [("cmu_tables", "cmu_step"), ("synth_tables", "synth_step")]


def _tables(g, name):
    return co.Tables(**{k: g[name][k] for k in co.TABLE_KEYS})


def _fields(state):
    st = state[:, :24]
    return st[..., 0:3], st[..., 3:7], st[..., 7:10], st[..., 10:13]


@pytest.mark.parametrize("tn,sn", CASES)
def test_frame_blend_bit_exact(golden, tn, sn):
    T, S = golden[tn], golden[sn]
    ids = S["in_motion_ids"]
    for tag in ("t0", "t1"):
        i0, i1, bl = co.frame_blend(S[tag], T["motion_len"][ids], T["num_frames"][ids], T["motion_dt"][ids])
        assert_equal(i0, S[f"{tag}_idx0"], f"{sn} {tag} idx0")
        assert_equal(i1, S[f"{tag}_idx1"], f"{sn} {tag} idx1")
        assert_equal(bl.view(np.uint32), S[f"{tag}_blend"].view(np.uint32), f"{sn} {tag} blend bits")


@pytest.mark.parametrize("tn,sn", CASES)
def test_motion_state(golden, tn, sn):
    S = golden[sn]
    tab = _tables(golden, tn)
    for tag in ("t0", "t1"):
        out = co.motion_state(tab, S["in_motion_ids"], S[tag], S["in_global_offset"])
        for k in co.STATE_KEYS:
            assert_close(out[k], S[f"{tag}_{k}"], what=f"{sn} {tag} {k}")
    out = co.motion_state(tab, S["in_motion_ids"], S["t0"], None)
    assert_close(out["rg_pos"], S["t0_rg_pos_nooffset"], what="rg_pos without offset")
    assert_close(out["root_pos"], S["t0_root_pos_smpl"], what="get_root_pos_smpl")
    # lerped quantities use only mul/add: they must agree to the bit with torch CPU
    out = co.motion_state(tab, S["in_motion_ids"], S["t0"], S["in_global_offset"])
    for k in ("rg_pos", "body_vel", "body_ang_vel", "dof_vel"):
        assert_equal(out[k].view(np.uint32), S[f"t0_{k}"].reshape(out[k].shape).view(np.uint32), f"{k} bits")


@pytest.mark.parametrize("tn,sn", CASES)
def test_functions_on_reference_states(golden, tn, sn):
    """Each envs/common.py function on the reference's own motion-state outputs."""
    S = golden[sn]
    bp, br, bv, ba = _fields(S["in_body_state"])
    r0 = [S[f"t0_{k}"] for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel")]
    r1 = [S[f"t1_{k}"] for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel")]
    rew, raw = co.imitation_reward(bp, br, bv, ba, *r0, K, W)
    assert_close(rew, S["reward_nopower"], what="reward")
    assert_close(raw, S["reward_raw4"], what="reward_raw")
    obs_self = co.self_obs(bp, br, bv, ba)
    assert_close(obs_self, S["obs"][:, :358], what="self obs", row_scale=True)
    obs_task = co.imitation_obs_v6(bp[:, 0], br[:, 0], bp, br, bv, ba, *r1)
    assert_close(obs_task, S["obs"][:, 358:], what="task obs", row_scale=True)
    n = 32
    assert_close(co.self_obs(bp[:n], br[:n], bv[:n], ba[:n], local_root_obs=False, root_height_obs=False, upright=False),
                 S["self_obs_variant"], what="self obs variant flags")
    assert_close(co.imitation_obs_v6(bp[:n, 0], br[:n, 0], bp[:n], br[:n], bv[:n], ba[:n], *[x[:n] for x in r1], upright=False),
                 S["task_obs_notupright"], what="task obs not upright")
    prog, pt = S["in_progress"], S["pass_time"]
    rs, tm = co.im_reset(prog, bp, r0[0], pt, True, np.full(24, 0.25, np.float32), False)
    assert_equal(rs, S["reset_train"], "reset train")
    assert_equal(tm, S["terminated_train"], "terminated train")
    rs, tm = co.im_reset(prog, bp[:, EVAL_IDS], r0[0][:, EVAL_IDS], pt, True, np.full(20, 0.5, np.float32), True)
    assert_equal(rs, S["reset_eval"], "reset eval")
    assert_equal(tm, S["terminated_eval"], "terminated eval")
    rs, tm = co.im_reset(prog, bp, r0[0], pt, False, np.full(24, 0.25, np.float32), False)
    assert_equal(rs, S["reset_noearly"], "reset no early termination")
    assert_equal(tm, S["terminated_noearly"], "terminated no early termination")


@pytest.mark.parametrize("tn,sn", CASES)
def test_full_step(golden, tn, sn):
    S = golden[sn]
    tab = _tables(golden, tn)
    args = (tab, S["in_body_state"], S["in_progress"], S["in_start_time"], S["in_start_offset"], S["in_motion_ids"],
            S["in_global_offset"], 1.0 / 30.0, K, W)
    r = co.step(*args, np.full(24, 0.25, np.float32), dof_force=S["in_dof_force"], dof_vel=S["in_dof_vel"], want_ref=True)
    assert_close(r["obs"], S["obs"], what="obs", row_scale=True)
    assert_close(r["reward"], S["reward"], what="reward")
    assert_close(r["reward_raw"], S["reward_raw"], what="reward_raw")
    assert_equal(r["reset"], S["reset_train"], "reset")
    assert_equal(r["terminated"], S["terminated_train"], "terminated")
    assert_close(r["ref_t"][:, :72], S["t0_rg_pos"].reshape(-1, 72), what="ref pos t")
    assert_close(r["ref_t1"][:, 72:168], S["t1_rb_rot"].reshape(-1, 96), what="ref rot t+1")
    r = co.step(*args, np.full(24, 0.5, np.float32), reset_body_mask=EVAL_MASK, use_mean=True)
    assert_equal(r["reset"], S["reset_eval"], "reset eval")
    assert_equal(r["terminated"], S["terminated_eval"], "terminated eval")
    assert_close(r["reward"], S["reward_nopower"], what="reward without power term")
    # the fixtures must exercise both outcomes of every flag
    for k in ("reset_train", "terminated_train", "reset_eval", "pass_time"):
        assert 0 < S[k].sum() < S[k].size, k


def test_gae_bit_exact(golden):
    G = golden["gae"]
    for tag in "abcde":
        gam, lam = G[f"{tag}_gamma_lambda"]
        adv = co.gae(G[f"{tag}_dones"], G[f"{tag}_values"], G[f"{tag}_rewards"], gam, lam)
        assert_equal(adv.view(np.uint32), G[f"{tag}_adv"].view(np.uint32), f"gae case {tag}")
    assert G["a_adv"][-1] == 0.0


def test_rms(golden):
    R, x1, x2 = golden["rms"], golden["cmu_step"]["obs"], golden["synth_step"]["obs"][:200]
    m, v, c = co.rms_update(x1, np.zeros((1, 934)), np.ones((1, 934)), np.ones(1))
    assert_close(m, R["mean1"], what="mean after 1 update")
    assert_close(v, R["var1"], rtol=1e-5, atol=1e-9, what="var after 1 update")
    assert_equal(c, R["count1"], "count")
    m, v, c = co.rms_update(x2, m, v, c)
    assert_close(m, R["mean2"], what="mean after 2 updates")
    assert_close(v, R["var2"], rtol=1e-5, atol=1e-9, what="var after 2 updates")
    assert_equal(c, R["count2"], "count")
    # not bit-exact by construction: torch's vectorised CPU sqrt is 1 ulp off the correctly rounded result
    # for ~0.6 % of inputs (measured in the build container), the oracle uses IEEE sqrtf like torch-CUDA.
    y = co.rms_forward(R["fwd_in"], R["mean2"], R["var2"])
    assert_close(y, R["fwd_out"], rtol=1e-6, atol=0, what="forward")
    assert y.max() == 10.0 and y.min() == -10.0
    y = co.rms_forward(R["fwd_in"], np.zeros(934), np.ones(934))
    assert_close(y, R["fwd_out_fresh"], rtol=1e-6, atol=0, what="forward, fresh state")


def test_sample_time_interval(golden):
    S, T = golden["sample_time"], golden["synth_tables"]
    t = co.sample_time_interval(S["phase"], T["motion_len"][S["ids"]], div_mode=0)
    assert_equal(t.view(np.uint32), S["time_interval"].view(np.uint32), "sample_time_interval bits (CPU semantics)")
    t = co.sample_time_interval(S["phase_trunc"], T["motion_len"][S["ids"]] - np.float32(0.1), div_mode=0)
    assert_equal(t.view(np.uint32), S["time_interval_trunc"].view(np.uint32), "truncate_time variant")


def test_amp_observations(golden_amp):
    A = golden_amp
    base = (A["root_pos"], A["root_rot"], A["root_vel"], A["root_ang_vel"], A["dof_pos"], A["dof_vel"], A["key_pos"])
    for name in ("default", "full_dof", "global_root", "not_upright", "with_params"):
        lro, rho, sub, shp, limb, up = (bool(x) for x in A[f"flags_{name}"])
        got = co.amp_obs(*base, A["subset"] if sub else None, lro, rho, up)
        want = A[f"amp_{name}"]
        assert_close(got, want[:, : got.shape[1]], what=f"amp obs {name}", row_scale=True)
        if shp:      # pass-through columns
            assert_equal(want[:, got.shape[1]:], np.concatenate([A["shape"], A["limb"]], -1), "shape/limb columns")


def test_mpjpe(golden):
    """extras["mpjpe"] of the evaluation step (humanoid_phc.py:159-163) against the reference's torch expression."""
    from conftest import load_npz
    M = load_npz("mpjpe.npz")
    for sn in ("cmu_step", "synth_step"):
        S = golden[sn]
        got = co.mpjpe(S["in_body_state"][:, :24, 0:3], S["t0_rg_pos"])
        assert_close(got, M[sn], what=f"mpjpe {sn}")
