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
