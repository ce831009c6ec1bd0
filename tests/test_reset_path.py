"""Reset path against a replay of the reference flow with the C oracle."""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_equal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_reset_envs_matches_reference_flow(golden):
    from oracle import c_oracle as co
    from puffer_phc_b200 import synth
    from puffer_phc_b200.envs.reset import EnvTensors, auto_reset, reset_envs
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    T = golden["synth_tables"]
    lib = MotionLibSMPL.from_tables({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in T.items()}, device=DEV)
    tab = co.Tables(**{k: T[k] for k in co.TABLE_KEYS})
    tdev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in T.items()}
    N = 200
    S = synth.make_env_state(tdev, N, seed=5, bodies_per_env=25)
    g = torch.Generator().manual_seed(3)
    reset_buf = (torch.rand(N, generator=g) < 0.3).to(DEV)
    term = reset_buf & (torch.rand(N, generator=g) < 0.5).to(DEV)
    env = EnvTensors(rigid_body_state=S["body_state"].clone(), humanoid_root_states=torch.zeros(N, 13, device=DEV),
                     dof_pos=torch.zeros(N, 69, device=DEV), dof_vel=torch.zeros(N, 69, device=DEV), progress_buf=S["progress"].clone(),
                     reset_buf=reset_buf.clone(), terminate_buf=term.clone(), global_offset=S["global_offset"].clone(),
                     motion_start_times=S["start_time"].clone(), motion_start_times_offset=S["start_offset"].clone() + 0.01,
                     sampled_motion_ids=S["motion_ids"].clone(), obs_buf=torch.full((N, 934), -7.0, device=DEV))
    before = {k: getattr(env, k).clone() for k in ("rigid_body_state", "humanoid_root_states", "dof_pos", "progress_buf", "obs_buf")}
    env_ids = torch.nonzero(reset_buf).squeeze(-1)
    # the same RNG call the reference makes
    torch.manual_seed(11)
    times = reset_envs(env, lib, env_ids)
    torch.manual_seed(11)
    ids = S["motion_ids"][env_ids]
    phase = torch.rand(ids.shape, device=DEV)
    want_times = ((phase * lib._motion_lengths[ids]) / (1 / 30)).long() * (1 / 30)
    assert torch.equal(times, want_times)
    torch.cuda.synchronize()
    # replay with the oracle
    e = env_ids.cpu().numpy()
    ms = co.motion_state(tab, S["motion_ids"].cpu().numpy()[e], times.cpu().numpy(), S["global_offset"].cpu().numpy()[e])
    bs = env.rigid_body_state.cpu().numpy()
    assert_close(bs[e][:, :24, 0:3], ms["rg_pos"], what="rigid body pos")
    assert_close(bs[e][:, :24, 3:7], ms["rb_rot"], what="rigid body rot")
    assert_close(bs[e][:, :24, 7:10], ms["body_vel"], what="rigid body vel")
    assert_close(bs[e][:, :24, 10:13], ms["body_ang_vel"], what="rigid body ang vel")
    root = env.humanoid_root_states.cpu().numpy()[e]
    assert_close(root, np.concatenate([ms["root_pos"], ms["root_rot"], ms["root_vel"], ms["root_ang_vel"]], -1), what="root states")
    assert_close(env.dof_pos.cpu().numpy()[e], ms["dof_pos"], what="dof_pos")
    assert_close(env.dof_vel.cpu().numpy()[e], ms["dof_vel"], what="dof_vel")
    keep = (~reset_buf).cpu().numpy()
    assert_equal(bs[keep], before["rigid_body_state"].cpu().numpy()[keep], "untouched envs keep their state")
    assert_equal(bs[e][:, 24:], before["rigid_body_state"].cpu().numpy()[e][:, 24:], "extra actor bodies untouched")
    assert_equal(env.obs_buf.cpu().numpy()[keep], before["obs_buf"].cpu().numpy()[keep], "untouched envs keep their obs")
    assert float(env.global_offset[env_ids].abs().sum()) == 0 and int(env.progress_buf[env_ids].abs().sum()) == 0
    assert not bool(env.reset_buf.any()) and not bool(env.terminate_buf[env_ids].any())
    assert torch.equal(env.motion_start_times[env_ids], times) and float(env.motion_start_times_offset[env_ids].abs().sum()) == 0
    # observation of the reset envs: body at the new state, reference at (0 + 1) * dt + start, zero offset
    t1 = (np.float32(1) * np.float32(1 / 30) + times.cpu().numpy()) + np.float32(0)
    ref1 = co.motion_state(tab, S["motion_ids"].cpu().numpy()[e], t1, np.zeros((len(e), 3), np.float32))
    st = bs[e][:, :24]
    want_obs = np.concatenate([co.self_obs(st[..., 0:3], st[..., 3:7], st[..., 7:10], st[..., 10:13]),
                               co.imitation_obs_v6(st[:, 0, 0:3], st[:, 0, 3:7], st[..., 0:3], st[..., 3:7], st[..., 7:10], st[..., 10:13],
                                                   ref1["rg_pos"], ref1["rb_rot"], ref1["body_vel"], ref1["body_ang_vel"])], -1)
    assert_close(env.obs_buf.cpu().numpy()[e], want_obs, what="obs of the reset envs", row_scale=True)

    # auto-reset bookkeeping (clean_pufferl/env.py:111-140)
    env.reset_buf[:] = reset_buf
    env.terminate_buf[:] = term
    terminals, truncs, masks = (torch.zeros(N, dtype=torch.bool, device=DEV) for _ in range(3))
    ep_ret, ep_len = torch.arange(N, device=DEV, dtype=torch.float32), torch.ones(N, device=DEV)
    rew = torch.full((N,), 0.5, device=DEV)
    idx, fin_ret, fin_len = auto_reset(env, lib, terminals, truncs, masks, ep_ret, ep_len, rew)
    assert torch.equal(idx, env_ids) and torch.equal(terminals, term)
    assert torch.equal(truncs, reset_buf & ~term) and torch.equal(masks, ~(reset_buf & ~term))
    assert torch.equal(fin_ret, torch.arange(N, device=DEV, dtype=torch.float32)[env_ids])
    want_ret = torch.where(reset_buf, torch.zeros(N, device=DEV), torch.arange(N, device=DEV, dtype=torch.float32)) + 0.5
    assert torch.equal(ep_ret, want_ret) and torch.equal(ep_len, torch.where(reset_buf, torch.zeros(N, device=DEV), torch.ones(N, device=DEV)) + 1)
    empty = reset_envs(env, lib, env_ids[:0])
    assert empty.numel() == 0


def _make_env(S, N, reset_buf, term, obs=None):
    from puffer_phc_b200.envs.reset import EnvTensors
    return EnvTensors(rigid_body_state=S["body_state"].clone(), humanoid_root_states=torch.zeros(N, 13, device=DEV),
                      dof_pos=torch.zeros(N, 69, device=DEV), dof_vel=torch.zeros(N, 69, device=DEV), progress_buf=S["progress"].clone(),
                      reset_buf=reset_buf.clone(), terminate_buf=term.clone(), global_offset=S["global_offset"].clone(),
                      motion_start_times=S["start_time"].clone(), motion_start_times_offset=S["start_offset"].clone() + 0.01,
                      sampled_motion_ids=S["motion_ids"].clone(), obs_buf=torch.full((N, 934), -7.0, device=DEV) if obs is None else obs.clone())


@pytest.mark.parametrize("N,bodies,random_start", [(200, 25, True), (3000, 24, True), (1, 24, True), (777, 24, False)])
def test_fused_auto_reset_equals_the_eager_reference_flow(golden, N, bodies, random_start):
    """phc_auto_reset (two launches, no host sync) against the eager path above (itself pinned to the oracle replay): every env
    tensor bit-identical, terminals / truncations / masks, episode returns / lengths, the ordered id list and the metric sums."""
    from puffer_phc_b200 import synth
    from puffer_phc_b200.envs.reset import AutoReset, auto_reset
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    T = golden["synth_tables"]
    tdev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in T.items()}
    lib = MotionLibSMPL.from_tables(tdev, device=DEV)
    S = synth.make_env_state(tdev, N, seed=9, bodies_per_env=bodies)
    g = torch.Generator().manual_seed(4)
    reset_buf = (torch.rand(N, generator=g) < 0.3).to(DEV)
    term = reset_buf & (torch.rand(N, generator=g) < 0.5).to(DEV)
    phase = torch.rand(N, generator=g).to(DEV)
    rew = torch.rand(N, generator=g).to(DEV)
    raw = torch.rand(N, 5, generator=g).to(DEV)
    ret0 = (torch.rand(N, generator=g) * 10).to(DEV)
    len0 = torch.randint(0, 300, (N,), generator=g).to(DEV)
    # eager reference flow
    A = _make_env(S, N, reset_buf, term)
    ids = torch.nonzero(reset_buf).squeeze(-1)
    K = int(ids.numel())
    times = lib.time_interval_from_phase(phase[:K], lib._motion_lengths[S["motion_ids"][ids]]) if random_start else torch.zeros(K, device=DEV)
    tA, trA, mA = (torch.zeros(N, dtype=torch.bool, device=DEV) for _ in range(3))
    retA, lenA = ret0.clone(), len0.clone().float()
    auto_reset(A, lib, tA, trA, mA, retA, lenA, rew, motion_times=times if K else None)
    # fused
    B = _make_env(S, N, reset_buf, term)
    ar = AutoReset(B, lib, random_start=random_start, ref_device="cpu")
    ar.episode_returns.copy_(ret0)
    ar.episode_lengths.copy_(len0.to(torch.int32))
    tB, trB, mB = ar(rew, raw, phase)
    torch.cuda.synchronize()
    for name in ("rigid_body_state", "humanoid_root_states", "dof_pos", "dof_vel", "progress_buf", "reset_buf", "terminate_buf",
                 "global_offset", "motion_start_times", "motion_start_times_offset", "obs_buf"):
        a, b = getattr(A, name).cpu().numpy(), getattr(B, name).cpu().numpy()
        if a.dtype == np.float32:
            a, b = a.view(np.uint32), b.view(np.uint32)
        assert_equal(b, a, f"env.{name}")
    assert_equal(tB.cpu().numpy(), tA.cpu().numpy(), "terminals")
    assert_equal(trB.cpu().numpy(), trA.cpu().numpy(), "truncations")
    assert_equal(mB.cpu().numpy(), mA.cpu().numpy(), "masks")
    assert_equal(ar.episode_returns.cpu().numpy().view(np.uint32), retA.cpu().numpy().view(np.uint32), "episode_returns")
    assert_equal(ar.episode_lengths.cpu().numpy(), lenA.cpu().numpy().astype(np.int32), "episode_lengths")
    assert int(ar.reset_count) == K
    assert_equal(ar.reset_ids[:K].cpu().numpy(), ids.cpu().numpy(), "ordered reset ids")
    m = ar.metric_values()
    rb, tb = reset_buf.cpu().numpy(), term.cpu().numpy()
    want = {"steps": N, "reward": rew.double().sum().item(), "resets": rb.sum(), "terminations": tb.sum(), "truncations": (rb & ~tb).sum(),
            "episodes": rb.sum(), "episode_return": ret0.double().cpu().numpy()[rb].sum(), "episode_length": len0.double().cpu().numpy()[rb].sum(),
            "r_pos": raw[:, 0].double().sum().item(), "r_power": raw[:, 4].double().sum().item()}
    for k, v in want.items():
        assert abs(m[k] - float(v)) <= 1e-9 * max(1.0, abs(float(v))), (k, m[k], v)
    # a second call with nothing flagged touches nothing but the episode counters
    before = B.obs_buf.clone()
    ar(rew, raw, phase)
    assert int(ar.reset_count) == 0 and torch.equal(B.obs_buf, before) and bool(ar.masks.all()) and not bool(ar.terminals.any())


def test_step_plus_auto_reset_in_one_cuda_graph_and_corrected_moments(golden):
    """FusedStep + AutoReset captured in ONE CUDA graph (no host sync anywhere); the RunningNorm moments accumulated by the step over
    pre-reset observations are corrected to the rows the reference stores (post-reset row of a terminated env, no row of a truncated
    env), the normalised copy of the reset rows is refreshed, and the episode metrics ride in the same statistics buffer."""
    from oracle import c_oracle as co
    from puffer_phc_b200 import synth
    from puffer_phc_b200.envs.reset import AutoReset, EnvTensors
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    from puffer_phc_b200.policies.running_norm import RunningNorm
    T = golden["synth_tables"]
    tdev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in T.items()}
    lib = MotionLibSMPL.from_tables(tdev, device=DEV)
    N = 1500
    S = synth.make_env_state(tdev, N, seed=21)
    rms = RunningNorm(934).to(DEV)
    rms.running_mean.uniform_(-0.2, 0.2)
    rms.running_var.uniform_(0.5, 2.0)
    fs = FusedStep(lib, N, StepConfig(ref_device="cpu"), rms=rms, normalize=True, accumulate_moments=True, defer_moments=True, metrics=True)
    env = EnvTensors(rigid_body_state=S["body_state"], humanoid_root_states=torch.zeros(N, 13, device=DEV), dof_pos=torch.zeros(N, 69, device=DEV),
                     dof_vel=S["dof_vel"], progress_buf=S["progress"], reset_buf=fs.reset_buf, terminate_buf=fs.terminate_buf,
                     global_offset=S["global_offset"], motion_start_times=S["start_time"], motion_start_times_offset=S["start_offset"],
                     sampled_motion_ids=S["motion_ids"], obs_buf=fs.obs_buf)
    ar = AutoReset(env, lib, ref_device="cpu", obs_norm=fs.obs_norm, rms=rms, fused=fs)
    phase = torch.rand(N, generator=torch.Generator().manual_seed(8)).to(DEV)
    saved = {k: v.clone() for k, v in S.items()}

    def body():
        fs(S["body_state"], S["progress"], S["start_time"], S["start_offset"], S["motion_ids"], S["global_offset"], S["dof_force"], S["dof_vel"])
        ar(fs.rew_buf, fs.reward_raw, phase)

    # eager run first (the expected result), then restore the inputs and replay the captured graph
    body()
    fs.flush_moments()
    torch.cuda.synchronize()
    want = {"obs": fs.obs_buf.clone(), "obs_norm": fs.obs_norm.clone(), "stats": fs.stats.clone(), "bs": S["body_state"].clone(),
            "masks": ar.masks.clone(), "terminals": ar.terminals.clone(), "ret": ar.episode_returns.clone()}
    n_reset = int(ar.reset_count)
    assert 0 < n_reset < N
    # moments over the stored rows, fp64 numpy: every env contributes its FINAL obs row unless it was truncated
    keep = ar.masks.cpu().numpy()
    X = want["obs"].cpu().numpy().astype(np.float64)[keep]
    st = want["stats"].cpu().numpy()
    assert st[0] == keep.sum()
    np.testing.assert_allclose(st[1:935], X.sum(0), rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(st[935:1869], (X * X).sum(0), rtol=1e-11, atol=1e-9)
    mv = fs.metric_values()
    assert mv["steps"] == N and mv["resets"] == n_reset and mv["truncations"] == (~keep).sum() and mv["episodes"] == n_reset
    # the normalised copy follows the final rows
    yn = co.rms_forward(want["obs"].cpu().numpy(), rms.running_mean.cpu().numpy(), rms.running_var.cpu().numpy())
    assert_close(want["obs_norm"].cpu().numpy(), yn, rtol=1e-5, atol=2e-6, what="obs_norm after the auto-reset")
    # ---- the same in one CUDA graph ----
    for k, v in saved.items():
        S[k].copy_(v)
    fs.stats.zero_()
    ar.episode_returns.zero_(); ar.episode_lengths.zero_()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body()                                  # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for k, v in saved.items():
        S[k].copy_(v)
    fs.partials.zero_(); fs.metric_partials.zero_(); fs.row_adjust.zero_(); fs.stats.zero_(); fs._pending_rows = 0
    ar.episode_returns.zero_(); ar.episode_lengths.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    graph.replay()
    fs.count_replayed(1)
    fs.flush_moments()
    torch.cuda.synchronize()
    assert torch.equal(fs.obs_buf, want["obs"]) and torch.equal(fs.obs_norm, want["obs_norm"]) and torch.equal(S["body_state"], want["bs"])
    assert torch.equal(ar.masks, want["masks"]) and torch.equal(ar.terminals, want["terminals"]) and torch.equal(ar.episode_returns, want["ret"])
    assert torch.equal(fs.stats, want["stats"])
