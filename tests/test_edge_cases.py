"""Empty, single-element and degenerate inputs through the drop-in surface (GPU).  What the reference's torch code does on the same
inputs is stated next to each case; c_gae.pyx on an empty array returns an empty array, on one element [0]."""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_equal, load_npz

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K = dict(k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1, w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def lib():
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    T = load_npz("synth_tables.npz")
    return MotionLibSMPL.from_tables({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in T.items()}, device=DEV)


def test_empty_batches(lib):
    """torch returns empty tensors of the right shape for empty batches; so do the kernels (no launch at all)."""
    from puffer_phc_b200 import c_gae
    from puffer_phc_b200.envs import common
    from puffer_phc_b200.policies.running_norm import RunningNorm
    st = lib.get_motion_state(torch.zeros(0, dtype=torch.long, device=DEV), torch.zeros(0, device=DEV))
    assert st["rg_pos"].shape == (0, 24, 3) and st["dof_pos"].shape == (0, 69) and st["motion_bodies"].shape == (0, 17)
    z3, z4 = torch.zeros(0, 24, 3, device=DEV), torch.zeros(0, 24, 4, device=DEV)
    obs = common.compute_imitation_observations_v6(z3[:, 0], z4[:, 0], z3, z4, z3, z3, z3, z4, z3, z3, 1, True)
    assert obs.shape == (0, 576)
    assert common.compute_humanoid_observations_smpl_max(z3, z4, z3, z3, None, None, True, True, True, False, False).shape == (0, 358)
    rew, raw = common.compute_imitation_reward(z3[:, 0], z4[:, 0], z3, z4, z3, z3, z3, z4, z3, z3, K)
    assert rew.shape == (0,) and raw.shape == (0, 4)
    reset, term = common.compute_humanoid_im_reset(torch.zeros(0, dtype=torch.bool, device=DEV), torch.zeros(0, dtype=torch.int16, device=DEV),
                                                   None, None, z3, z3, torch.zeros(0, dtype=torch.bool, device=DEV), True,
                                                   torch.full((24,), 0.25, device=DEV), False)
    assert reset.shape == (0,) and term.shape == (0,) and reset.dtype == torch.bool
    assert c_gae.compute_gae_cuda(torch.zeros(0, device=DEV), torch.zeros(0, device=DEV), torch.zeros(0, device=DEV), 0.98, 0.2).shape == (0,)
    assert c_gae.compute_gae(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32), 0.98, 0.2).shape == (0,)
    assert RunningNorm(934).to(DEV)(torch.zeros(0, 934, device=DEV)).shape == (0, 934)


def test_single_elements(lib):
    from oracle import c_oracle as co
    from puffer_phc_b200 import c_gae
    # c_gae.pyx: a single element gets advantage 0 (c_gae.pyx:21-31: the loop body never runs)
    a = c_gae.compute_gae(np.array([0.0], np.float32), np.array([1.5], np.float32), np.array([2.0], np.float32), 0.98, 0.2)
    assert a.tolist() == [0.0]
    a = c_gae.compute_gae(np.array([0.0, 1.0], np.float32), np.array([1.5, -0.5], np.float32), np.array([2.0, 3.0], np.float32), 0.98, 0.2)
    assert_equal(a.view(np.uint32), co.gae(np.array([0.0, 1.0]), np.array([1.5, -0.5]), np.array([2.0, 3.0]), 0.98, 0.2).view(np.uint32), "L=2")
    # one query, at the very start, at the very end and far beyond the end of the clip (frame index saturates, blend = 1)
    T = load_npz("synth_tables.npz")
    tab = co.Tables(**{k: T[k] for k in co.TABLE_KEYS})
    L = float(T["motion_len"][2])
    for t in (0.0, L, L + 7.0, -3.0):
        got = lib.get_motion_state(torch.tensor([2], device=DEV), torch.tensor([t], device=DEV), debug=True)
        want, (i0, i1, bl) = co.motion_state(tab, np.array([2]), np.array([t], np.float32), debug=True)
        assert int(got["frame_idx0"][0]) == int(i0[0]) and int(got["frame_idx1"][0]) == int(i1[0])
        assert float(got["blend"][0]) == float(bl[0])
        for k in ("rg_pos", "rb_rot", "dof_pos", "body_vel"):
            assert_close(got[k].cpu().numpy(), want[k], what=f"t={t} {k}")
    nf = int(T["num_frames"][2])
    got = lib.get_motion_state(torch.tensor([2], device=DEV), torch.tensor([L + 7.0], device=DEV), debug=True)
    assert int(got["frame_idx0"][0]) == int(got["frame_idx1"][0]) == nf - 1 and float(got["blend"][0]) == 1.0


def test_nan_sim_state_does_not_terminate_or_crash(lib):
    """A NaN simulator state (PhysX blow-up): every comparison with NaN is False in torch, so the env is NOT marked fallen;
    rewards / observations carry the NaN.  The fused step must behave the same and must not disturb the other envs."""
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    S = load_npz("synth_step.npz")
    args = [cu(S[k]) for k in ("in_body_state", "in_progress", "in_start_time", "in_start_offset", "in_motion_ids", "in_global_offset",
                               "in_dof_force", "in_dof_vel")]
    N = args[1].shape[0]
    fs = FusedStep(lib, N, StepConfig())
    clean = {k: v.clone() for k, v in fs(*args).items()}
    bad = args[0].clone()
    bad[5] = float("nan")
    out = fs(bad, *args[1:])
    assert bool(torch.isnan(out["reward"][5])) and bool(torch.isnan(out["obs"][5]).any())
    assert not bool(out["terminated"][5])
    assert bool(out["reset"][5]) == bool(cu(S["in_progress"])[5].float() * (1 / 30) + cu(S["in_start_time"])[5] + cu(S["in_start_offset"])[5]
                                        >= lib._motion_lengths[cu(S["in_motion_ids"])[5]])
    keep = torch.ones(N, dtype=torch.bool, device=DEV)
    keep[5] = False
    for k in ("obs", "reward", "reward_raw", "reset", "terminated"):
        assert torch.equal(out[k][keep], clean[k][keep]), k


def test_calc_frame_blend_entry_point_bit_exact(lib):
    """MotionLibBase._calc_frame_blend (row A1) with the reference's signature, against the reference's own outputs."""
    T, S = load_npz("synth_tables.npz"), load_npz("synth_step.npz")
    ids = S["in_motion_ids"]
    for tag in ("t0", "t1"):
        i0, i1, bl = lib._calc_frame_blend(cu(S[tag]), cu(T["motion_len"][ids]), cu(T["num_frames"][ids]), cu(T["motion_dt"][ids]))
        assert i0.dtype == torch.int64 and i1.dtype == torch.int64 and bl.dtype == torch.float32
        assert_equal(i0.cpu().numpy(), S[f"{tag}_idx0"], f"{tag} idx0")
        assert_equal(i1.cpu().numpy(), S[f"{tag}_idx1"], f"{tag} idx1")
        assert_equal(bl.cpu().numpy().view(np.uint32), S[f"{tag}_blend"].view(np.uint32), f"{tag} blend bits")
    e = torch.zeros(0, device=DEV)
    assert lib._calc_frame_blend(e, e, torch.zeros(0, dtype=torch.long, device=DEV), e)[0].shape == (0,)
    with pytest.raises(ValueError):
        lib._calc_frame_blend(torch.zeros(3, device=DEV), torch.zeros(2, device=DEV), torch.zeros(3, dtype=torch.long, device=DEV), torch.zeros(3, device=DEV))


def test_sampling_weight_updates():
    """update_hard / update_soft_sampling_weight / update_sampling_prob (motion_lib.py:454-508) drive load_motions' multinomial."""
    from types import SimpleNamespace
    from puffer_phc_b200.motion_file import RawClips
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    g = load_npz("loader.npz")
    raw = RawClips.from_dict({f"clip{i}": {"root_trans_offset": g[f"clip{i}_root_trans_offset"], "pose_aa": g[f"clip{i}_pose_aa"],
                                          "pose_quat_global": g[f"clip{i}_pose_quat_global"], "beta": np.zeros(16), "fps": int(g[f"clip{i}_fps"])}
                              for i in range(4)})
    ml = MotionLibSMPL(SimpleNamespace(motion_file=raw, device=DEV, min_length=-1, max_length=50, im_eval=False, is_deterministic=False, step_dt=1 / 30))
    assert ml._sampling_prob.tolist() == [0.25] * 4
    ml.update_hard_sampling_weight(["clip1", "clip3"])
    assert ml._sampling_prob.tolist() == [0.0, 0.5, 0.0, 0.5]
    torch.manual_seed(0)
    from puffer_phc_b200.skeleton import SkeletonTree
    sk = SkeletonTree([f"b{j}" for j in range(24)], g["parents"].astype(np.int32), g["local_translation"])
    ml.load_motions(skeleton_trees=[sk] * 16, gender_betas=torch.zeros(16, 17), limb_weights=np.zeros((16, 10)))     # random_sample=True
    assert set(ml._curr_motion_ids.tolist()) <= {1, 3} and len(ml.curr_motion_keys) == 16
    assert ml.sample_motions(8).max() < 16
    ml.update_hard_sampling_weight([])
    assert ml._sampling_prob.tolist() == [0.25] * 4
    ml.update_soft_sampling_weight(["clip0"])
    ml.update_soft_sampling_weight(["clip0", "clip2"])
    assert ml._termination_history.tolist() == [2.0, 0.0, 1.0, 0.0]
    assert_close(ml._sampling_prob.cpu().numpy(), np.array([2 / 3, 0, 1 / 3, 0], np.float32), rtol=1e-6, atol=0, what="soft prob")
    assert ml.update_sampling_prob(torch.zeros(4, device=DEV)) is False and ml.update_sampling_prob(torch.ones(3, device=DEV)) is False
    assert ml._get_num_bodies() == 24
