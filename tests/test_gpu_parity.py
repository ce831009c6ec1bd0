"""Parity of the CUDA kernels (called through the C ABI via the drop-in Python surface) against
(1) the golden vectors produced by the reference itself and (2) the C oracle on seeded inputs.
Integers / flags: bit-exact.  Floats: |a-b| <= 1e-5*|b| + 2e-6 (north_star tolerance; conftest.RTOL/ATOL)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_equal

pytestmark = pytest.mark.gpu

K = dict(k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1, w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1)
KL, WL = [100.0, 10.0, 0.1, 0.1], [0.5, 0.3, 0.1, 0.1]
EVAL_IDS = [j for j in range(24) if j not in (4, 8, 18, 23)]
CASES = [("cmu_tables", "cmu_step"), ("synth_tables", "synth_step")]
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def make_lib(T, pack=True):
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    return MotionLibSMPL.from_tables({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in T.items()}, device=DEV, pack=pack)


def fields(state):
    st = state[:, :24]
    return st[..., 0:3], st[..., 3:7], st[..., 7:10], st[..., 10:13]


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tn,sn", CASES)
def test_motion_state_vs_golden(golden, tn, sn):
    from puffer_phc_b200.motion_lib import STATE_KEYS
    S = golden[sn]
    lib = make_lib(golden[tn])
    ids, off = cu(S["in_motion_ids"]), cu(S["in_global_offset"])
    for tag in ("t0", "t1"):
        out = lib.get_motion_state(ids, cu(S[tag]), off, debug=True)
        assert set(STATE_KEYS) <= set(out)
        assert_equal(npy(out["frame_idx0"]), S[f"{tag}_idx0"], f"{tag} idx0")
        assert_equal(npy(out["frame_idx1"]), S[f"{tag}_idx1"], f"{tag} idx1")
        assert_equal(npy(out["blend"]).view(np.uint32), S[f"{tag}_blend"].view(np.uint32), f"{tag} blend bits")
        for k in STATE_KEYS:
            assert tuple(out[k].shape) == S[f"{tag}_{k}"].shape, k
            assert_close(npy(out[k]), S[f"{tag}_{k}"], what=f"{sn} {tag} {k}")
        # lerp outputs are pure mul/add chains -> identical bits
        for k in ("rg_pos", "body_vel", "body_ang_vel", "dof_vel", "motion_aa"):
            assert_equal(npy(out[k]).view(np.uint32), S[f"{tag}_{k}"].view(np.uint32), f"{k} bits")
    out = lib.get_motion_state(ids, cu(S["t0"]), None)
    assert_close(npy(out["rg_pos"]), S["t0_rg_pos_nooffset"], what="no offset")
    out = lib.get_root_pos_smpl(ids, cu(S["t0"]))
    assert list(out) == ["root_pos"]
    assert_close(npy(out["root_pos"]), S["t0_root_pos_smpl"], what="get_root_pos_smpl")
    assert_equal(npy(lib.get_motion_length(ids)), golden[tn]["motion_len"][S["in_motion_ids"]], "get_motion_length")
    empty = lib.get_motion_state(ids[:0], cu(S["t0"])[:0], off[:0])
    assert empty["rg_pos"].shape == (0, 24, 3)


@pytest.mark.parametrize("tn,sn", CASES)
def test_common_functions_vs_golden(golden, tn, sn):
    from puffer_phc_b200.envs import common
    S = golden[sn]
    state = cu(S["in_body_state"])                       # [N, bodies, 13]: the functions get strided views of it
    bp, br, bv, ba = fields(state)
    assert not bp.is_contiguous()
    r0 = [cu(S[f"t0_{k}"]) for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel")]
    r1 = [cu(S[f"t1_{k}"]) for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel")]
    rew, raw = common.compute_imitation_reward(bp[:, 0], br[:, 0], bp, br, bv, ba, *r0, K)
    assert_close(npy(rew), S["reward_nopower"], what="reward")
    assert_close(npy(raw), S["reward_raw4"], what="reward_raw")
    obs_self = common.compute_humanoid_observations_smpl_max(bp, br, bv, ba, None, None, True, True, True, False, False)
    assert_close(npy(obs_self), S["obs"][:, :358], what="self obs", row_scale=True)
    obs_task = common.compute_imitation_observations_v6(bp[:, 0], br[:, 0], bp, br, bv, ba, *r1, 1, True)
    assert_close(npy(obs_task), S["obs"][:, 358:], what="task obs", row_scale=True)
    # contiguous copies give the same bits as the strided views
    obs_task_c = common.compute_imitation_observations_v6(bp[:, 0].contiguous(), br[:, 0].contiguous(), bp.contiguous(), br.contiguous(),
                                                          bv.contiguous(), ba.contiguous(), *r1, 1, True)
    assert torch.equal(obs_task, obs_task_c)
    n = 32
    v = common.compute_humanoid_observations_smpl_max(bp[:n], br[:n], bv[:n], ba[:n], None, None, False, False, False, False, False)
    assert_close(npy(v), S["self_obs_variant"], what="self obs flags variant", row_scale=True)
    v = common.compute_imitation_observations_v6(bp[:n, 0], br[:n, 0], bp[:n], br[:n], bv[:n], ba[:n], *[x[:n] for x in r1], 1, False)
    assert_close(npy(v), S["task_obs_notupright"], what="task obs upright=False", row_scale=True)
    prog, pt = cu(S["in_progress"]), cu(S["pass_time"])
    rb = torch.ones(len(prog), dtype=torch.bool, device=DEV)
    contact, cids = torch.zeros(len(prog), 24, 3, device=DEV), torch.zeros(4, dtype=torch.long, device=DEV)
    rs, tm = common.compute_humanoid_im_reset(rb, prog, contact, cids, bp.clone(), r0[0].clone(), pt, True,
                                              torch.full((24,), 0.25, device=DEV), False)
    assert rs.dtype == torch.bool and tm.dtype == torch.bool
    assert_equal(npy(rs), S["reset_train"], "reset train")
    assert_equal(npy(tm), S["terminated_train"], "terminated train")
    rs, tm = common.compute_humanoid_im_reset(rb, prog, contact, cids, bp[..., EVAL_IDS, :].clone(), r0[0][..., EVAL_IDS, :].clone(), pt,
                                              True, torch.full((24,), 0.5, device=DEV)[..., EVAL_IDS], True)
    assert_equal(npy(rs), S["reset_eval"], "reset eval")
    assert_equal(npy(tm), S["terminated_eval"], "terminated eval")
    rs, tm = common.compute_humanoid_im_reset(rb, prog, contact, cids, bp, r0[0], pt, False, torch.full((24,), 0.25, device=DEV), False)
    assert_equal(npy(rs), S["reset_noearly"], "reset without early termination")
    assert_equal(npy(tm), S["terminated_noearly"], "terminated without early termination")


def _fused(lib, S, cfg=None, power=True, **kw):
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    cfg = cfg or StepConfig(use_power_reward=power)
    N = S["in_progress"].shape[0]
    fs = FusedStep(lib, N, cfg, **kw)
    out = fs(cu(S["in_body_state"]), cu(S["in_progress"]), cu(S["in_start_time"]), cu(S["in_start_offset"]), cu(S["in_motion_ids"]),
             cu(S["in_global_offset"]), cu(S["in_dof_force"]) if cfg.use_power_reward else None,
             cu(S["in_dof_vel"]) if cfg.use_power_reward else None)
    torch.cuda.synchronize()
    return fs, out


@pytest.mark.parametrize("pack", [True, False])
@pytest.mark.parametrize("tn,sn", CASES)
def test_fused_step_vs_golden(golden, tn, sn, pack):
    from puffer_phc_b200.fused_step import StepConfig
    S = golden[sn]
    lib = make_lib(golden[tn], pack=pack)
    fs, out = _fused(lib, S, debug_ref=True)
    assert_close(npy(out["obs"]), S["obs"], what="obs", row_scale=True)
    assert_close(npy(out["reward"]), S["reward"], what="reward")
    assert_close(npy(out["reward_raw"]), S["reward_raw"], what="reward_raw")
    assert_equal(npy(out["reset"]), S["reset_train"], "reset")
    assert_equal(npy(out["terminated"]), S["terminated_train"], "terminated")
    assert_close(npy(fs.ref_t)[:, :72], S["t0_rg_pos"].reshape(-1, 72), what="ref pos at t")
    assert_close(npy(fs.ref_t1)[:, 72:168], S["t1_rb_rot"].reshape(-1, 96), what="ref rot at t+1")
    fs, out = _fused(lib, S, cfg=StepConfig(use_power_reward=False).eval_mode())
    assert_equal(npy(out["reset"]), S["reset_eval"], "reset eval")
    assert_equal(npy(out["terminated"]), S["terminated_eval"], "terminated eval")
    assert_close(npy(out["reward"]), S["reward_nopower"], what="reward without power")
    assert out["reward_raw"].shape[1] == 4
    fs, out = _fused(lib, S, cfg=StepConfig(enable_early_termination=False))
    assert_equal(npy(out["reset"]), S["reset_noearly"], "reset without early termination")


def test_fused_step_ragged_and_layouts(golden):
    """N not a multiple of 8, N=1, unaligned sim records and a user-supplied pitched obs buffer give the same rows."""
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    S = golden["synth_step"]
    lib = make_lib(golden["synth_tables"])
    _, full = _fused(lib, S)
    full = {k: v.clone() for k, v in full.items()}
    for n in (1, 7, 13, 250):
        sub = {k: (v[:n] if k.startswith("in_") else v) for k, v in S.items()}
        _, out = _fused(lib, sub)
        for k in ("obs", "reward", "reward_raw", "reset", "terminated"):
            assert torch.equal(out[k], full[k][:n]), (n, k)
    # unaligned PhysX stride (25 bodies -> 325 floats, not a multiple of 4) exercises the scalar staging path
    st = torch.zeros(256, 25, 13, device=DEV)
    st[:, :24] = cu(S["in_body_state"])[:, :24]
    sub = dict(S)
    fs = FusedStep(lib, 256, StepConfig())
    out = fs(st, cu(S["in_progress"]), cu(S["in_start_time"]), cu(S["in_start_offset"]), cu(S["in_motion_ids"]), cu(S["in_global_offset"]),
             cu(S["in_dof_force"]), cu(S["in_dof_vel"]))
    assert torch.equal(out["obs"], full["obs"])
    # pitched output rows (936 floats)
    pitched = torch.zeros(256, 936, device=DEV)
    out = fs(st, cu(S["in_progress"]), cu(S["in_start_time"]), cu(S["in_start_offset"]), cu(S["in_motion_ids"]), cu(S["in_global_offset"]),
             cu(S["in_dof_force"]), cu(S["in_dof_vel"]), out={"obs": pitched[:, :934]})
    assert torch.equal(pitched[:, :934], full["obs"]) and float(pitched[:, 934:].abs().sum()) == 0.0


def test_fused_step_rms_outputs(golden):
    """In-kernel RunningNorm.forward and the fp64 column moments against the oracle."""
    from oracle import c_oracle as co
    from puffer_phc_b200.policies.running_norm import RunningNorm
    S, R = golden["cmu_step"], golden["rms"]
    lib = make_lib(golden["cmu_tables"])
    rms = RunningNorm(934).to(DEV)
    rms.running_mean.copy_(cu(R["mean2"]))
    rms.running_var.copy_(cu(R["var2"]))
    fs, out = _fused(lib, S, rms=rms, normalize=True, accumulate_moments=True)
    got_obs = npy(out["obs"])
    want = co.rms_forward(got_obs, R["mean2"], R["var2"])
    # the fused kernel multiplies by an IEEE reciprocal of sqrt(var+eps) (one per column) instead of dividing: <= 1.5 ulp
    assert_close(npy(out["obs_norm"]), want, rtol=5e-7, atol=1e-7, what="normalised obs vs forward() of the same obs")
    assert_close(npy(out["obs_norm"]), npy(rms(out["obs"])), rtol=5e-7, atol=1e-7, what="fused vs stand-alone forward kernel")
    assert_close(npy(out["obs_norm"]), co.rms_forward(S["obs"], R["mean2"], R["var2"]), rtol=1e-4, atol=1e-4, what="normalised obs vs reference obs", row_scale=True)
    m = npy(rms.moments_buffer())
    obs64 = got_obs.astype(np.float64)
    assert m[0] == 256
    np.testing.assert_allclose(m[1:935], obs64.sum(0), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(m[935:], (obs64 ** 2).sum(0), rtol=1e-12, atol=1e-12)
    rms2 = RunningNorm(934).to(DEV)
    rms2.moments_buffer().copy_(rms.moments_buffer())
    rms2.finalize()
    wm, wv, wc = co.rms_update(got_obs, np.zeros((1, 934)), np.ones((1, 934)), np.ones(1))
    assert_close(npy(rms2.running_mean), wm, what="running_mean from in-kernel moments")
    assert_close(npy(rms2.running_var), wv, rtol=1e-5, atol=1e-9, what="running_var from in-kernel moments")
    assert float(rms2.count) == 2.0 and float(rms2.moments_buffer().abs().sum()) == 0.0


def test_running_norm_vs_golden(golden):
    from puffer_phc_b200.policies.running_norm import RunningNorm
    R, x1, x2 = golden["rms"], golden["cmu_step"]["obs"], golden["synth_step"]["obs"][:200]
    rn = RunningNorm(934).to(DEV)
    fresh = rn(cu(R["fwd_in"]))
    assert_close(npy(fresh), R["fwd_out_fresh"], rtol=1e-6, atol=0, what="forward with the initial state")
    rn.update(cu(x1))
    assert_close(npy(rn.running_mean), R["mean1"], what="mean after 1 update")
    assert_close(npy(rn.running_var), R["var1"], rtol=1e-5, atol=1e-9, what="var after 1 update")
    assert_equal(npy(rn.count), R["count1"], "count")
    rn.update(cu(x2))
    assert_close(npy(rn.running_mean), R["mean2"], what="mean after 2 updates")
    assert_close(npy(rn.running_var), R["var2"], rtol=1e-5, atol=1e-9, what="var after 2 updates")
    assert_equal(npy(rn.count), R["count2"], "count")
    rn.running_mean.copy_(cu(R["mean2"])); rn.running_var.copy_(cu(R["var2"]))
    y = rn(cu(R["fwd_in"]))
    assert_close(npy(y), R["fwd_out"], rtol=1e-6, atol=0, what="forward")
    assert float(y.max()) == 10.0 and float(y.min()) == -10.0
    # strided rows and 3-D input
    wide = torch.zeros(16, 940, device=DEV)
    wide[:, :934] = cu(R["fwd_in"])
    assert torch.equal(rn(wide[:, :934]), y)
    assert torch.equal(rn(cu(R["fwd_in"]).view(4, 4, 934)).view(16, 934), y)
    sd = rn.state_dict()
    assert set(sd) == {"running_mean", "running_var", "count"} and sd["running_mean"].shape == (1, 934)
    import pickle
    rn2 = pickle.loads(pickle.dumps(rn))
    assert torch.equal(rn2.running_mean, rn.running_mean) and rn2.clip == rn.clip


def test_gae_vs_golden_and_reference(golden):
    from puffer_phc_b200 import c_gae
    G = golden["gae"]
    for tag in "abcde":
        gam, lam = (float(x) for x in G[f"{tag}_gamma_lambda"])
        for mode in (0, 2):
            adv = c_gae.compute_gae_cuda(cu(G[f"{tag}_dones"]), cu(G[f"{tag}_values"]), cu(G[f"{tag}_rewards"]), gam, lam, mode=mode)
            assert_equal(npy(adv).view(np.uint32), G[f"{tag}_adv"].view(np.uint32), f"gae case {tag} mode {mode}")
    # numpy in -> numpy out, like the reference module
    adv = c_gae.compute_gae(G["a_dones"], G["a_values"], G["a_rewards"], 0.98, 0.2)
    assert isinstance(adv, np.ndarray) and adv.dtype == np.float32
    assert_equal(adv.view(np.uint32), G["a_adv"].view(np.uint32), "numpy path")
    assert adv[-1] == 0.0


@pytest.mark.parametrize("gam,lam", [(0.98, 0.2), (0.99, 0.95), (0.9, 0.5), (1.0, 1.0), (0.0, 0.5)])
def test_gae_full_size_vs_oracle(gam, lam):
    """BASELINE config 3 (4096 x 32) and a ragged length against the C oracle (and the reference's own compiled
    c_gae when oracle/_ref holds it): bit-exact."""
    from oracle import c_oracle as co
    from puffer_phc_b200 import c_gae, synth
    for n_env, hor in ((4096, 32), (37, 53)):
        R = synth.make_rollout(n_env, hor, seed=2, p_done=0.01 if gam < 1.0 else 0.05)
        d, v, r = (R[k].numpy() for k in ("dones", "values", "rewards"))
        want = co.gae(d, v, r, gam, lam)
        got = npy(c_gae.compute_gae_cuda(cu(d), cu(v), cu(r), gam, lam))
        assert_equal(got.view(np.uint32), want.view(np.uint32), f"gae {n_env}x{hor} gamma={gam} lambda={lam}")
        try:
            import sys, os
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref"))
            import c_gae as ref_gae
        except ImportError:
            continue
        assert_equal(got.view(np.uint32), ref_gae.compute_gae(d, v, r, gam, lam).view(np.uint32), "vs reference c_gae")


@pytest.mark.parametrize("gam,lam", [(0.98, 0.2), (0.9, 0.5), (0.0, 0.5)])
def test_gae_short_window_many_tiles_ragged_and_unaligned(gam, lam):
    """The short-window kernel on the big-tile path (65536 x 32) and the small-tile path, lengths that end inside a tile / a chunk / a 16-byte
    group, and array slices that are not 16-byte aligned (4-byte staging, scalar stores): bit-exact against the serial scan."""
    from oracle import c_oracle as co
    from puffer_phc_b200 import c_gae, synth
    R = synth.make_rollout(65536, 32, seed=11, p_done=0.01)
    d, v, r = (R[k].numpy() for k in ("dones", "values", "rewards"))
    dc, vc, rc = cu(d), cu(v), cu(r)
    for L in (65536 * 32, 65536 * 32 - 3, 2048 * 296 + 1, 2048 * 296, 2048 * 300 + 2047, 511, 513, 9, 2, 1):
        want = co.gae(d[:L], v[:L], r[:L], gam, lam)
        got = npy(c_gae.compute_gae_cuda(dc[:L], vc[:L], rc[:L], gam, lam))
        assert_equal(got.view(np.uint32), want.view(np.uint32), f"gae L={L} gamma={gam} lambda={lam}")
    for off, L in ((1, 700001), (3, 4097), (2, 700)):
        want = co.gae(d[off:off + L], v[off:off + L], r[off:off + L], gam, lam)
        out = torch.full((L + 8,), -7.0, device=DEV)
        got = c_gae.compute_gae_cuda(dc[off:off + L], vc[off:off + L], rc[off:off + L], gam, lam, out=out[off:off + L])
        assert_equal(npy(got).view(np.uint32), want.view(np.uint32), f"gae unaligned off={off} L={L}")
        assert torch.all(out[:off] == -7.0) and torch.all(out[off + L:] == -7.0), "stores outside the output slice"


def test_sample_time_interval(golden):
    S, T = golden["sample_time"], golden["synth_tables"]
    lib = make_lib(T)
    t = lib.time_interval_from_phase(cu(S["phase"]), cu(T["motion_len"][S["ids"]]), cpu_division=True)
    assert_equal(npy(t).view(np.uint32), S["time_interval"].view(np.uint32), "CPU-reference semantics")
    # default = what the reference computes when it runs on CUDA tensors (motion_lib.py:526-535 evaluated by torch-CUDA)
    phase, ln = cu(S["phase"]), cu(T["motion_len"][S["ids"]])
    curr_fps = 1 / 30
    want = ((phase * ln) / curr_fps).long() * curr_fps
    got = lib.time_interval_from_phase(phase, ln)
    assert torch.equal(got, want)
    torch.manual_seed(5)
    a = lib.sample_time_interval(cu(S["ids"]))
    torch.manual_seed(5)
    ph = torch.rand(S["ids"].shape, device=DEV)
    assert torch.equal(a, ((ph * ln) / curr_fps).long() * curr_fps)
    assert_equal(npy(lib.get_motion_num_steps()), S["num_steps_all"], "get_motion_num_steps")
    torch.manual_seed(7)
    ids = lib.sample_motions(1000)
    torch.manual_seed(7)
    assert torch.equal(ids, torch.multinomial(lib._sampling_batch_prob, num_samples=1000, replacement=True))


@pytest.mark.parametrize("upright", [True, False])
def test_imitation_obs_future_steps(golden, upright):
    """time_steps = 3 (common.py:106-176 with its .view(B, time_steps, J, .) of the reference tensors): every future step is the
    single-step observation of the same simulated bodies against that step's reference -- checked against the C oracle step by step
    and, when oracle/_ref is staged, against the reference's own function on torch-CUDA."""
    from oracle import c_oracle as co
    from puffer_phc_b200.envs import common
    S = golden["synth_step"]
    st = cu(S["in_body_state"])
    bp, br, bv, ba = fields(st)
    N, TS = bp.shape[0], 3
    g = torch.Generator().manual_seed(3)
    refs1 = [cu(S[f"t1_{k}"]) for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel")]
    # three different "future" references per env: the golden t+1 reference, and two perturbed copies (rotations stay unit)
    steps = []
    for s in range(TS):
        p = refs1[0] + 0.05 * s * torch.randn(refs1[0].shape, generator=g).to(DEV)
        q = refs1[1].roll(s, 0).contiguous()
        v = refs1[2] * (1.0 + 0.1 * s)
        w = refs1[3].roll(-s, 0).contiguous()
        steps.append((p, q, v, w))
    stacked = [torch.stack([steps[s][k] for s in range(TS)], 1).contiguous() for k in range(4)]          # [N, TS, J, k]
    got = common.compute_imitation_observations_v6(bp[:, 0], br[:, 0], bp, br, bv, ba, *stacked, TS, upright)
    assert got.shape == (N, TS * 24 * 24)
    for s in range(TS):
        want = co.imitation_obs_v6(npy(bp[:, 0]), npy(br[:, 0]), npy(bp), npy(br), npy(bv), npy(ba), *[npy(x) for x in steps[s]], 1, upright)
        assert_close(npy(got[:, s * 576:(s + 1) * 576]), want, what=f"future step {s}", row_scale=True)
        one = common.compute_imitation_observations_v6(bp[:, 0], br[:, 0], bp, br, bv, ba, *steps[s], 1, upright)
        assert torch.equal(got[:, s * 576:(s + 1) * 576], one), "a future step differs from the single-step kernel on the same inputs"
    from oracle import ref_runner as rr
    if rr.available():
        Rm = rr.boot()
        ref = Rm.common.compute_imitation_observations_v6(bp[:, 0].contiguous(), br[:, 0].contiguous(), bp.contiguous(), br.contiguous(),
                                                          bv.contiguous(), ba.contiguous(), *stacked, TS, upright)
        assert_close(npy(got), npy(ref), what="vs the reference's own function, time_steps = 3", row_scale=True)


def test_error_behaviour(golden):
    from puffer_phc_b200.envs import common
    from puffer_phc_b200 import c_gae
    S = golden["synth_step"]
    bp, br, bv, ba = fields(cu(S["in_body_state"]))
    r1 = [cu(S[f"t1_{k}"]) for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel")]
    with pytest.raises(ValueError):
        common.compute_imitation_observations_v6(bp[:, 0], br[:, 0], bp, br, bv, ba, *r1, 0, True)
    with pytest.raises(RuntimeError):
        common.compute_imitation_observations_v6(bp[:, 0].cpu(), br[:, 0], bp, br, bv, ba, *r1, 1, True)
    with pytest.raises(ValueError):
        c_gae.compute_gae_cuda(torch.zeros(4, device=DEV), torch.zeros(5, device=DEV), torch.zeros(4, device=DEV), 0.9, 0.9)
    lib = make_lib(golden["synth_tables"])
    with pytest.raises(RuntimeError):
        lib.get_motion_state(torch.zeros(2, dtype=torch.long), torch.zeros(2), None)


# ---- BASELINE configs 1, 2 and 4 at full size against the C oracle ------------------------------------
@pytest.fixture(scope="module")
def amass_lib():
    from puffer_phc_b200 import synth
    T = synth.make_motion_library(11313, seed=0, device=DEV)
    host = {k: v.cpu().numpy() for k, v in T.items()}
    return T, host


def _check_against_oracle(T_host, lib, N, seed, tables_dev):
    from oracle import c_oracle as co
    from puffer_phc_b200 import synth
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    S = synth.make_env_state(tables_dev, N, seed=seed)
    fs = FusedStep(lib, N, StepConfig())
    out = fs(S["body_state"], S["progress"], S["start_time"], S["start_offset"], S["motion_ids"], S["global_offset"], S["dof_force"], S["dof_vel"])
    torch.cuda.synchronize()
    tab = co.Tables(**{k: T_host[k] for k in co.TABLE_KEYS})
    want = co.step(tab, npy(S["body_state"]), npy(S["progress"]), npy(S["start_time"]), npy(S["start_offset"]), npy(S["motion_ids"]),
                   npy(S["global_offset"]), 1.0 / 30.0, KL, WL, np.full(24, 0.25, np.float32), dof_force=npy(S["dof_force"]),
                   dof_vel=npy(S["dof_vel"]))
    assert_equal(npy(out["reset"]), want["reset"], f"N={N} reset")
    assert_equal(npy(out["terminated"]), want["terminated"], f"N={N} terminated")
    assert 0 < want["reset"].sum() < N and 0 < want["terminated"].sum() < N
    assert_close(npy(out["obs"]), want["obs"], what=f"N={N} obs", row_scale=True)
    assert_close(npy(out["reward"]), want["reward"], what=f"N={N} reward")
    assert_close(npy(out["reward_raw"]), want["reward_raw"], what=f"N={N} reward_raw")
    # size-independent properties: heading-frame rotation preserves lengths; obs blocks are consistent
    obs = out["obs"]
    st = S["body_state"][:, :24]
    local = obs[:, 1:70].view(N, 23, 3).norm(dim=-1)
    world = (st[:, 1:, 0:3] - st[:, :1, 0:3]).norm(dim=-1)
    assert torch.allclose(local, world, rtol=1e-4, atol=1e-5)
    assert torch.equal(obs[:, 0], st[:, 0, 2])
    tn = obs[:, 70:214].view(N, 24, 2, 3)
    assert torch.allclose(tn.norm(dim=-1), torch.ones_like(tn[..., 0]), atol=1e-3)
    return S, out


def test_config1_cmu_1024_envs(golden):
    """BASELINE config 1: the real clip, 1024 envs, fused step vs the C oracle."""
    T = golden["cmu_tables"]
    lib = make_lib(T)
    tables_dev = {k: cu(v) for k, v in T.items()}
    _check_against_oracle(T, lib, 1024, 1, tables_dev)


def test_config2_amass_4096_envs(amass_lib):
    """BASELINE config 2: 4096 envs over the 11313-clip synthetic library."""
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    T, host = amass_lib
    lib = MotionLibSMPL.from_tables(T, device=DEV)
    _check_against_oracle(host, lib, 4096, 1, T)


@pytest.mark.parametrize("n", [148 * 3 + 1, 148 * 5 + 3, 148 * 8 - 1, 148 * 12 + 5, 148 * 12 * 2 + 11])
def test_balanced_blocks_all_widths(amass_lib, n):
    """The launcher spreads small batches evenly over the CTAs' iterations (here: 4, 6, 8 envs per block in one iteration, 8 in two,
    10 in three, each with a ragged last block): same flags and rows as the oracle whatever the block width."""
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    T, host = amass_lib
    lib = MotionLibSMPL.from_tables(T, device=DEV)
    _check_against_oracle(host, lib, n, 5, T)


def test_config4_amass_65536_envs(amass_lib):
    """BASELINE config 4 (one rank's share): 65536 envs; also the stand-alone query against the fused one."""
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    T, host = amass_lib
    lib = MotionLibSMPL.from_tables(T, device=DEV)
    S, out = _check_against_oracle(host, lib, 65536, 3, T)
    # motion-state drop-in at full size vs oracle
    from oracle import c_oracle as co
    tab = co.Tables(**{k: host[k] for k in co.TABLE_KEYS})
    t0 = (S["progress"].float() * torch.tensor(1.0 / 30.0, device=DEV) + S["start_time"]) + S["start_offset"]
    got = lib.get_motion_state(S["motion_ids"], t0, S["global_offset"], debug=True)
    want, (i0, i1, bl) = co.motion_state(tab, npy(S["motion_ids"]), npy(t0), npy(S["global_offset"]), debug=True)
    assert_equal(npy(got["frame_idx0"]), i0, "idx0")
    assert_equal(npy(got["frame_idx1"]), i1, "idx1")
    assert_equal(npy(got["blend"]).view(np.uint32), bl.view(np.uint32), "blend bits")
    for k in want:
        assert_close(npy(got[k]), want[k], what=f"65536 {k}")


def test_fused_step_cuda_graph_replay(golden):
    """The step captured in a CUDA graph (small batches are launch-bound) reproduces the eager outputs bit for bit and
    follows in-place updates of the input buffers."""
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    S = golden["synth_step"]
    lib = make_lib(golden["synth_tables"])
    _, eager = _fused(lib, S)
    eager = {k: v.clone() for k, v in eager.items()}
    N = S["in_progress"].shape[0]
    fs = FusedStep(lib, N, StepConfig())
    bufs = [cu(S[k]) for k in ("in_body_state", "in_progress", "in_start_time", "in_start_offset", "in_motion_ids", "in_global_offset",
                               "in_dof_force", "in_dof_vel")]
    graph, out = fs.capture(*bufs)
    out["obs"].zero_()
    graph.replay()
    torch.cuda.synchronize()
    for k in ("obs", "reward", "reward_raw", "reset", "terminated"):
        assert torch.equal(out[k], eager[k]), k
    bufs[1].add_(1)                       # progress_buf += 1, in place, as the env does before the next step
    graph.replay()
    want = fs(*bufs, out={k: torch.empty_like(v) for k, v in eager.items()})
    torch.cuda.synchronize()
    assert torch.equal(out["obs"], want["obs"]) and not torch.equal(out["obs"], eager["obs"])


def test_running_norm_update_layouts():
    """Both moment kernels (streaming rows for even C, column chunks otherwise / for odd strides) against fp64 numpy."""
    from puffer_phc_b200.policies.running_norm import RunningNorm
    g = torch.Generator().manual_seed(0)
    for B, C, pitch in ((1000, 934, 934), (777, 935, 935), (513, 934, 937), (3, 6, 6), (70000, 934, 934)):
        x = (torch.randn(B, pitch, generator=g) * 3 + 1.5).to(DEV)[:, :C]
        rn = RunningNorm(C).to(DEV)
        rn.update(x)
        x64 = x.double().cpu().numpy()
        np.testing.assert_allclose(npy(rn.running_mean)[0], x64.mean(0), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(npy(rn.running_var)[0], x64.var(0), rtol=1e-5, atol=1e-9)
        assert float(rn.count) == 2.0


def test_amp_observations_vs_golden(golden_amp):
    """build_amp_observations_smpl / dof_to_obs_smpl (row f3) against the reference's outputs."""
    from puffer_phc_b200.envs import common
    A = golden_amp
    base = [cu(A[k]) for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "dof_pos", "dof_vel", "key_pos")]
    shape, limb, subset = cu(A["shape"]), cu(A["limb"]), cu(A["subset"])
    for name in ("default", "full_dof", "global_root", "not_upright", "with_params"):
        lro, rho, sub, shp, lw, up = (bool(x) for x in A[f"flags_{name}"])
        got = common.build_amp_observations_smpl(*base, shape, limb, subset, lro, rho, sub, shp, lw, up)
        assert tuple(got.shape) == A[f"amp_{name}"].shape
        assert_close(npy(got), A[f"amp_{name}"], what=f"amp obs {name}", row_scale=True)
    d = common.dof_to_obs_smpl(cu(A["dof_pos"])[:, A["subset"]])
    assert_close(npy(d), A["amp_default"][:, 13:13 + 114], what="dof_to_obs_smpl", row_scale=True)


def test_mpjpe_vs_golden(golden):
    """compute_mpjpe (evaluation metric, humanoid_phc.py:159-163) on the strided PhysX view against the reference's outputs."""
    from conftest import load_npz
    from puffer_phc_b200.envs import common
    M = load_npz("mpjpe.npz")
    for sn in ("cmu_step", "synth_step"):
        S = golden[sn]
        st = cu(S["in_body_state"])
        got = common.compute_mpjpe(st[:, :24, 0:3], cu(S["t0_rg_pos"]))        # non-contiguous view, consumed in place
        assert_close(npy(got), M[sn], what=f"mpjpe {sn}")
    assert common.compute_mpjpe(st[:0, :24, 0:3], cu(S["t0_rg_pos"])[:0]).shape == (0,)


def test_step_host_pipelined_equals_device_step(golden):
    """FusedStep.step_host (pinned host buffers, env ranges pipelined over copy / compute streams) returns exactly what the
    device-resident step returns, for any chunk count, and accumulates the same RunningNorm moments."""
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    from puffer_phc_b200.policies.running_norm import RunningNorm
    S = golden["synth_step"]
    lib = make_lib(golden["synth_tables"])
    N = S["in_progress"].shape[0]
    keys = {"body_state": "in_body_state", "progress": "in_progress", "start_time": "in_start_time", "start_offset": "in_start_offset",
            "motion_ids": "in_motion_ids", "global_offset": "in_global_offset", "dof_force": "in_dof_force", "dof_vel": "in_dof_vel"}
    host = {k: torch.from_numpy(np.ascontiguousarray(S[v])).pin_memory() for k, v in keys.items()}
    host["body_state"] = host["body_state"].reshape(N, -1)
    dev = {k: v.to(DEV) for k, v in host.items()}
    rms0 = RunningNorm(934).to(DEV)
    fs0 = FusedStep(lib, N, StepConfig(), rms=rms0, normalize=True, accumulate_moments=True)
    want = fs0(dev["body_state"], dev["progress"], dev["start_time"], dev["start_offset"], dev["motion_ids"], dev["global_offset"],
               dev["dof_force"], dev["dof_vel"])
    for chunks in (1, 3, 4, 1000):
        rms = RunningNorm(934).to(DEV)
        fs = FusedStep(lib, N, StepConfig(), rms=rms, normalize=True, accumulate_moments=True)
        got = fs.step_host(host, chunks=chunks)
        for k in ("reward", "reward_raw", "reset", "terminated"):
            assert torch.equal(got[k], want[k].cpu()), (chunks, k)
        assert torch.equal(fs.obs_buf, want["obs"]) and torch.equal(fs.obs_norm, want["obs_norm"]), chunks
        m, m0 = npy(rms.moments_buffer()), npy(rms0.moments_buffer())
        assert m[0] == m0[0] == N
        np.testing.assert_allclose(m, m0, rtol=1e-13, atol=1e-12)
    # two steps in flight (wait=False, double-buffered staging and result buffers): each handle returns its own step's results
    host2 = {k: v.clone().pin_memory() for k, v in host.items()}
    host2["progress"] = (host2["progress"] + 2).pin_memory()
    dev2 = {k: v.to(DEV) for k, v in host2.items()}
    want2 = {k: v.clone() for k, v in fs0(dev2["body_state"], dev2["progress"], dev2["start_time"], dev2["start_offset"], dev2["motion_ids"],
                                          dev2["global_offset"], dev2["dof_force"], dev2["dof_vel"]).items()}
    want1 = {k: v.clone() for k, v in fs0(dev["body_state"], dev["progress"], dev["start_time"], dev["start_offset"], dev["motion_ids"],
                                          dev["global_offset"], dev["dof_force"], dev["dof_vel"]).items()}
    fs = FusedStep(lib, N, StepConfig(), rms=RunningNorm(934).to(DEV), normalize=True, accumulate_moments=True)
    pending, seen = None, 0
    for i in range(6):
        h = fs.step_host(host2 if i % 2 else host, chunks=3, wait=False)
        if pending is not None:
            got, w = pending[0].result(), (want2 if pending[1] else want1)
            for k in ("reward", "reward_raw", "reset", "terminated"):
                assert torch.equal(got[k], w[k].cpu()), ("in flight", i, k)
            seen += 1
        pending = (h, i % 2)
    got = pending[0].result()
    assert torch.equal(got["reward"], want2["reward"].cpu()) and seen == 5
    torch.cuda.synchronize()
    assert torch.equal(fs.obs_buf, want2["obs"])
    with pytest.raises(ValueError):
        fs0(dev["body_state"], dev["progress"], dev["start_time"], dev["start_offset"], dev["motion_ids"], dev["global_offset"],
            dev["dof_force"], dev["dof_vel"], env_range=(4, 16))


def test_deferred_moments_equal_per_step_moments(golden):
    """defer_moments: the kernel adds each step's column sums to its per-CTA slots; one flush per rollout gives the same
    pending moments as folding after every step."""
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    from puffer_phc_b200.policies.running_norm import RunningNorm
    lib = make_lib(golden["synth_tables"])
    S = golden["synth_step"]
    args = [cu(S[k]) for k in ("in_body_state", "in_progress", "in_start_time", "in_start_offset", "in_motion_ids", "in_global_offset",
                               "in_dof_force", "in_dof_vel")]
    N = args[1].shape[0]
    ra, rb = RunningNorm(934).to(DEV), RunningNorm(934).to(DEV)
    fa = FusedStep(lib, N, StepConfig(), rms=ra, normalize=True, accumulate_moments=True)
    fb = FusedStep(lib, N, StepConfig(), rms=rb, normalize=True, accumulate_moments=True, defer_moments=True)
    for step in range(3):
        args[1] = args[1] + 1                      # progress_buf advances: different observations every step
        oa, ob = fa(*args), fb(*args)
        assert torch.equal(oa["obs"], ob["obs"]) and torch.equal(oa["obs_norm"], ob["obs_norm"])
    assert float(rb.moments_buffer().abs().sum()) == 0.0          # nothing folded yet
    fb.flush_moments()
    ma, mb = npy(ra.moments_buffer()), npy(rb.moments_buffer())
    assert ma[0] == mb[0] == 3 * N
    np.testing.assert_allclose(mb, ma, rtol=1e-13, atol=1e-10)
    fb.flush_moments()                                            # idempotent when nothing is pending
    np.testing.assert_allclose(npy(rb.moments_buffer()), mb, rtol=0, atol=0)
    ra.finalize(); rb.finalize()
    assert_close(npy(rb.running_mean), npy(ra.running_mean), rtol=1e-7, atol=1e-9, what="running_mean")
    assert_close(npy(rb.running_var), npy(ra.running_var), rtol=1e-7, atol=1e-12, what="running_var")


@pytest.mark.parametrize("flavour", ["cpu", "cuda"])
def test_pair_tables_do_not_change_the_fused_step(golden, flavour):
    """The motion library's pair tables (slerp quantities of every frame pair, blend == 0 fast path that skips frame 1) against the
    same kernel computing everything inline: identical outputs, bit for bit, up to the sign of exact zeros."""
    import ctypes as C
    from puffer_phc_b200 import _ffi, synth
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    T = synth.make_motion_library(300, seed=4, device=DEV, other_fps_fraction=0.3, freeze_every=5)
    lib = MotionLibSMPL.from_tables(T, device=DEV)
    N = 6001
    S = synth.make_env_state(T, N, seed=12)
    args = [S[k] for k in ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")]
    fs = FusedStep(lib, N, StepConfig(ref_device=flavour))
    with_aux = {k: v.clone() for k, v in fs(*args).items()}
    ct = lib.ctables_for(_ffi.ref_device(flavour))
    assert ct.pair_aux and ct.pair_flags and ct.pair_device == _ffi.ref_device(flavour)
    flags = lib._pair[_ffi.ref_device(flavour)][1]
    assert 0 < int(flags.sum()) < flags.numel()          # frozen runs take the midpoint fall-back, the rest does not
    plain = _ffi.MotionTables()
    C.memmove(C.byref(plain), C.byref(ct), C.sizeof(ct))
    plain.pair_aux, plain.pair_flags = None, None
    lib.ctables_for = lambda flavour_: plain              # the same kernel without the tables
    without = fs(*args)
    torch.cuda.synchronize()
    for k in ("obs", "reward", "reward_raw"):
        a, b = with_aux[k] + 0.0, without[k] + 0.0         # -0.0 + 0.0 = +0.0
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), k
    assert torch.equal(with_aux["reset"], without["reset"]) and torch.equal(with_aux["terminated"], without["terminated"])


def test_fused_stats_exchange_kernel_equals_reduce_plus_finalize(golden):
    """phc_stats_allreduce_finalize (fold partials + exchange + running-average update in one kernel; here with one rank) against
    phc_stats_reduce + phc_rms_finalize: running_mean / running_var / count and the metric sums bit-identical, slots cleared."""
    from puffer_phc_b200 import synth
    from puffer_phc_b200.dist import StatsExchange
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    from puffer_phc_b200.policies.running_norm import RunningNorm
    T = synth.make_motion_library(100, seed=2, device=DEV)
    lib = MotionLibSMPL.from_tables(T, device=DEV)
    N = 5000
    S = [synth.make_env_state(T, N, seed=40 + i) for i in range(3)]
    keys = ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")
    res = []
    for fused_exchange in (False, True):
        rms = RunningNorm(934).to(DEV)
        fs = FusedStep(lib, N, StepConfig(ref_device="cpu"), rms=rms, normalize=True, accumulate_moments=True, defer_moments=True, metrics=True)
        ex = StatsExchange(DEV) if fused_exchange else None
        for rollout in range(2):                    # two updates: the second one uses count = 2 and the first one's statistics
            for s in S:
                fs(*[s[k] for k in keys])
            if ex is not None:
                ex.allreduce_finalize(fs)
            else:
                fs.flush_moments()
                rms.finalize()
        torch.cuda.synchronize()
        assert float(fs.partials.abs().sum()) == 0 and float(fs.metric_partials.abs().sum()) == 0 and fs._pending_rows == 0
        res.append((rms.running_mean.clone(), rms.running_var.clone(), rms.count.clone(), fs.stats[1 + 2 * 934:].clone()))
    for a, b, what in zip(res[0], res[1], ("running_mean", "running_var", "count", "metric sums")):
        assert torch.equal(a, b), what
    assert float(res[1][2]) == 3.0 and float(res[1][3][0]) == 2 * 3 * N
