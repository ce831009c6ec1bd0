"""Row f4: the motion TABLE BUILD (reference MotionLibSMPL.load_motions / load_motion_with_skeleton,
puffer_phc/motion_lib.py:257-429, 744-825; SkeletonMotion.from_skeleton_state, poselib_skeleton.py:1167-1249).

CPU part: the C oracle (oracle/loader_oracle.c) against tables the reference's own loader built (tests/golden/loader.npz:
the real sample clip + three synthetic clips, deterministic and random-crop/random-heading loads), the flat file format,
the skeleton parser.  GPU part: ``MotionLibSMPL(cfg).load_motions`` (csrc/build_tables.cu through the C ABI) against the
same golden tables, against the oracle on a library with every tile-boundary length, and the de-duplicated layout.

Tolerances.  Positions, global/local rotations: BIT-EXACT (float32 forward kinematics in the reference's operation order).
Linear / angular velocity: bit-exact for the oracle; for the CUDA path 1e-6 relative because the float64 acos / exp of the
CUDA libm and glibc may differ in the last bit before the float32 rounding.  Dof velocity (float32 acos/sin/cos/atan2 from
three different libms): 1e-5 relative, the north-star tolerance.  Random heading (scipy's float64 Rotation algebra
restated): 1e-5 relative with a 1e-5 absolute floor, because a last-bit float64 difference can flip the float32 rounding of
a local rotation and the float32 forward kinematics amplifies that to ~1e-7 in positions, ~5e-6 in velocities.
"""
import os
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import assert_close, assert_equal, load_npz
from oracle import c_oracle as co

FIELDS = ("gts", "grs", "lrs", "gvs", "gavs", "dvs")
BITS = ("gts", "grs", "lrs")
HAS_CUDA = torch.cuda.is_available()
DEV = "cuda:0"


@pytest.fixture(scope="module")
def g():
    return load_npz("loader.npz")


def _clips(g):
    return [{"root_trans_offset": g[f"clip{i}_root_trans_offset"], "pose_aa": g[f"clip{i}_pose_aa"],
             "pose_quat_global": g[f"clip{i}_pose_quat_global"], "beta": np.zeros(16), "gender": "neutral",
             "fps": int(g[f"clip{i}_fps"])} for i in range(4)]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---- scipy's Rotation algebra for the random heading (motion_lib.py:789-799), float64 numpy restatement ----------------
def heading_rotate(q, trans, theta):
    """(h * Rotation.from_quat(q)).as_quat() and trans @ R(h)^T for h = rotation by theta about z."""
    s, c = np.sin(0.5 * theta), np.cos(0.5 * theta)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    r = np.stack([c * x - s * y, c * y + s * x, c * z + w * s, c * w - s * z], -1)
    r = r / np.linalg.norm(r, axis=-1, keepdims=True)
    R = np.array([[c * c - s * s, -2 * s * c, 0.0], [2 * s * c, c * c - s * s, 0.0], [0.0, 0.0, 1.0]])
    return r, trans @ R.T


def heading_rotvec(rv, theta):
    """(h * Rotation.from_rotvec(rv)).as_rotvec()"""
    s, c = np.sin(0.5 * theta), np.cos(0.5 * theta)
    ang = np.linalg.norm(rv, axis=-1)
    small = ang <= 1e-3
    scale = np.where(small, 0.5 - ang ** 2 / 48 + ang ** 4 / 3840, np.sin(ang / 2) / np.where(small, 1.0, ang))
    x, y, z, w = scale * rv[:, 0], scale * rv[:, 1], scale * rv[:, 2], np.cos(ang / 2)
    p = np.stack([c * x - s * y, c * y + s * x, c * z + w * s, c * w - s * z], -1)
    p = p / np.linalg.norm(p, axis=-1, keepdims=True)
    p = np.where(p[:, 3:] < 0, -p, p)
    ang = 2 * np.arctan2(np.linalg.norm(p[:, :3], axis=-1), p[:, 3])
    small = ang <= 1e-3
    scale = np.where(small, 2 + ang ** 2 / 12 + 7 * ang ** 4 / 2880, ang / np.sin(np.where(small, 1.0, ang) / 2))
    return scale[:, None] * p[:, :3]


def replay_rng(g, seed):
    """The reference's host RNG calls, in its order (motion_lib.py:773-791)."""
    random.seed(seed)
    np.random.seed(seed)
    cap, starts, headings = int(g["max_length"]), [], []
    for c in _clips(g):
        T = c["pose_quat_global"].shape[0]
        starts.append(random.randint(0, T - cap) if T >= cap else 0)
        headings.append(np.pi * (2 * np.random.random() - 1.0))
    return starts, headings


# ========================================================================================================================
# CPU: oracle pinned to the reference's loader
# ========================================================================================================================
def test_oracle_builds_reference_tables_bit_exact(g):
    cap = int(g["max_length"])
    for i, c in enumerate(_clips(g)):
        T = min(c["pose_quat_global"].shape[0], cap)                   # deterministic crop = [0, max_length)
        o = co.build_clip(c["pose_quat_global"][:T], c["root_trans_offset"][:T], g["parents"], g["local_translation"], c["fps"])
        a, n = int(g["tab_length_starts"][i]), int(g["tab_num_frames"][i])
        assert n == T
        for k in ("gts", "grs", "lrs", "gvs", "gavs"):
            assert_equal(_bits(o[k]), _bits(g["tab_" + k][a:a + n]), f"clip {i} {k} bits")
        assert_close(o["dvs"], g["tab_dvs"][a:a + n], rtol=1e-5, atol=1e-6, what=f"clip {i} dvs")


def test_oracle_random_crop_and_heading(g):
    starts, headings = replay_rng(g, int(g["rnd_seed"]))
    cap = int(g["max_length"])
    for i, c in enumerate(_clips(g)):
        a, n = int(g["rnd_length_starts"][i]), int(g["rnd_num_frames"][i])
        s = starts[i]
        q, tr = heading_rotate(c["pose_quat_global"][s:s + n], c["root_trans_offset"][s:s + n], headings[i])
        o = co.build_clip(q, tr, g["parents"], g["local_translation"], c["fps"])
        assert n == min(c["pose_quat_global"].shape[0], cap)
        for k in FIELDS:
            assert_close(o[k], g["rnd_" + k][a:a + n], rtol=1e-5, atol=1e-5, what=f"rnd clip {i} {k}")
    # _motion_aa: uncropped rows, root rotation vector rotated inside the crop only
    row = 0
    for i, c in enumerate(_clips(g)):
        T = c["pose_aa"].shape[0]
        n, s = int(g["rnd_num_frames"][i]), starts[i]
        want = g["rnd_motion_aa"][row:row + T]
        aa = c["pose_aa"].copy()
        aa[s:s + n, :3] = heading_rotvec(aa[s:s + n, :3], headings[i])
        assert_close(aa.astype(np.float32), want, rtol=1e-6, atol=1e-7, what=f"rnd clip {i} motion_aa")
        row += T


def test_oracle_rejects_one_frame_clip(g):
    c = _clips(g)[0]
    with pytest.raises(ValueError):
        co.build_clip(c["pose_quat_global"][:1], c["root_trans_offset"][:1], g["parents"], g["local_translation"], 30)


def test_flat_file_round_trip(tmp_path, g):
    from puffer_phc_b200.motion_file import RawClips
    raw = RawClips.from_dict({f"clip{i}": c for i, c in enumerate(_clips(g))})
    assert raw.num_frames.tolist() == [222, 60, 29, 33] and raw.fps.tolist() == [30, 30, 30, 60]
    path = str(tmp_path / "lib.phcmot")
    raw.save(path)
    back = RawClips.open(path)
    assert back.keys.tolist() == raw.keys.tolist() and back.gender == raw.gender
    for k in ("num_frames", "fps", "starts", "root_trans", "pose_aa", "pose_quat_global", "beta", "has_beta"):
        assert_equal(np.asarray(getattr(back, k)), np.asarray(getattr(raw, k)), k)
    sub = raw.subset([2, 0])
    assert sub.num_frames.tolist() == [29, 222]
    assert_equal(sub.clip(1)["pose_quat_global"], g["clip0_pose_quat_global"], "subset clip")
    with open(path, "r+b") as f:                                             # truncated / foreign files are refused
        f.truncate(os.path.getsize(path) - 8)
    with pytest.raises(ValueError):
        RawClips.load(path)


def test_skeleton_from_mjcf(tmp_path):
    from puffer_phc_b200.skeleton import SkeletonTree
    xml = """<mujoco><worldbody><body name="Pelvis" pos="0 0 1"><body name="L_Hip" pos="0.1 0.2 -0.3"><body name="L_Knee" pos="0 0 -0.4"/></body>
             <body name="R_Hip" pos="-0.1 0.2 -0.3"/></body></worldbody></mujoco>"""
    p = tmp_path / "h.xml"
    p.write_text(xml)
    t = SkeletonTree.from_mjcf(str(p))
    assert t.node_names == ["Pelvis", "L_Hip", "L_Knee", "R_Hip"]
    assert t.parent_indices.tolist() == [-1, 0, 1, 0]
    assert_equal(t.local_translation.numpy(), np.array([[0, 0, 1], [0.1, 0.2, -0.3], [0, 0, -0.4], [-0.1, 0.2, -0.3]], np.float32))
    ref_xml = "/root/reference/puffer_phc/assets/smpl_humanoid.xml"          # only in the build container
    if os.path.isfile(ref_xml):
        gz = load_npz("loader.npz")
        t = SkeletonTree.from_mjcf(ref_xml)
        assert_equal(t.parent_indices.numpy(), gz["parents"].astype(np.int32))
        assert_equal(t.local_translation.numpy(), gz["local_translation"])


def test_loader_refuses_cpu_device(g):
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    from puffer_phc_b200.motion_file import RawClips
    raw = RawClips.from_dict({f"clip{i}": c for i, c in enumerate(_clips(g))})
    cfg = SimpleNamespace(motion_file=raw, device="cpu", min_length=-1, max_length=50, im_eval=False, is_deterministic=True)
    with pytest.raises(RuntimeError):
        MotionLibSMPL(cfg)


# ========================================================================================================================
# GPU: csrc/build_tables.cu through MotionLibSMPL.load_motions
# ========================================================================================================================
def _skeleton(g):
    from puffer_phc_b200.skeleton import SkeletonTree
    return SkeletonTree([f"b{j}" for j in range(24)], g["parents"].astype(np.int32), g["local_translation"])


def _lib(raw, **kw):
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    cfg = SimpleNamespace(motion_file=raw, device=DEV, min_length=kw.pop("min_length", -1), max_length=kw.pop("max_length", 50),
                          im_eval=kw.pop("im_eval", False), is_deterministic=kw.pop("is_deterministic", True), step_dt=1 / 30,
                          num_thread=1)
    return MotionLibSMPL(cfg)


def _tables_np(lib):
    out = {k: getattr(lib, k).cpu().numpy() for k in FIELDS}
    out.update(motion_aa=lib._motion_aa.cpu().numpy(), motion_len=lib._motion_lengths.cpu().numpy(), motion_dt=lib._motion_dt.cpu().numpy(),
               motion_fps=lib._motion_fps.cpu().numpy(), num_frames=lib._motion_num_frames.cpu().numpy(),
               length_starts=lib.length_starts.cpu().numpy(), motion_bodies=lib._motion_bodies.cpu().numpy(),
               limb_weights=lib._motion_limb_weights.cpu().numpy())
    return out


@pytest.mark.gpu
def test_load_motions_matches_reference_loader(g):
    from puffer_phc_b200.motion_file import RawClips
    raw = RawClips.from_dict({f"clip{i}": c for i, c in enumerate(_clips(g))})
    lib = _lib(raw)
    sk = _skeleton(g)
    ret = lib.load_motions(skeleton_trees=[sk] * 4, gender_betas=torch.zeros(4, 17), limb_weights=np.zeros((4, 10)), random_sample=False)
    assert ret is None
    T = _tables_np(lib)
    for k in BITS:
        assert_equal(_bits(T[k]), _bits(g["tab_" + k]), f"{k} bits")
    for k in ("gvs", "gavs"):
        assert_close(T[k], g["tab_" + k], rtol=1e-6, atol=1e-7, what=k)
    assert_close(T["dvs"], g["tab_dvs"], rtol=1e-5, atol=1e-6, what="dvs")
    for k in ("motion_aa", "motion_len", "motion_dt", "motion_fps", "num_frames", "length_starts", "motion_bodies", "limb_weights"):
        assert_equal(T[k], g["tab_" + k], k)
    assert T["num_frames"].dtype == np.int64 and T["length_starts"].dtype == np.int64
    assert lib._curr_motion_ids.tolist() == [0, 1, 2, 3] and lib.curr_motion_keys.tolist() == [f"clip{i}" for i in range(4)]
    # the packed records written in the same pass == phc_pack_frames of the tables
    packed = lib.packed.clone()
    assert_equal(_bits(lib.pack().cpu().numpy()), _bits(packed.cpu().numpy()), "packed records")
    # and the library answers queries
    st = lib.get_motion_state(torch.tensor([0, 3], device=DEV), torch.tensor([0.5, 0.2], device=DEV))
    assert st["rg_pos"].shape == (2, 24, 3) and torch.isfinite(st["dof_pos"]).all()


@pytest.mark.gpu
def test_load_motions_random_crop_and_heading(g):
    from puffer_phc_b200.motion_file import RawClips
    raw = RawClips.from_dict({f"clip{i}": c for i, c in enumerate(_clips(g))})
    lib = _lib(raw, is_deterministic=False)
    seed = int(g["rnd_seed"])
    random.seed(seed)
    np.random.seed(seed)
    lib.load_motions(skeleton_trees=[_skeleton(g)] * 4, gender_betas=torch.zeros(4, 17), limb_weights=np.zeros((4, 10)),
                     random_sample=False, sample_idxes=torch.arange(4))
    starts, headings = replay_rng(g, seed)
    assert lib._crop_start.tolist() == starts and np.allclose(lib._heading, headings, rtol=0, atol=0)
    T = _tables_np(lib)
    for k in FIELDS:
        assert_close(T[k], g["rnd_" + k], rtol=1e-5, atol=1e-5, what="rnd " + k)
    assert_close(T["motion_aa"], g["rnd_motion_aa"], rtol=1e-6, atol=1e-7, what="rnd motion_aa")
    for k in ("motion_len", "motion_dt", "num_frames", "length_starts"):
        assert_equal(T[k], g["rnd_" + k], "rnd " + k)


def _synthetic_raw(lengths, seed=3, fps=(30, 60, 120)):
    """Clips of the given lengths: smooth random global rotations + a root walk, float64, a frozen run in every third clip."""
    from puffer_phc_b200.motion_file import RawClips
    rng = np.random.default_rng(seed)
    clips = {}
    for i, T in enumerate(lengths):
        t = np.arange(T)[:, None, None] / 30.0
        phase, freq = rng.uniform(0, 6.28, (1, 24, 3)), rng.uniform(0.2, 1.5, (1, 24, 3))
        ea = 0.7 * np.sin(freq * 6.28 * t + phase)                           # [T,24,3] axis-angle per body
        ang = np.linalg.norm(ea, axis=-1, keepdims=True)
        q = np.concatenate([ea / np.maximum(ang, 1e-12) * np.sin(ang / 2), np.cos(ang / 2)], -1)
        q *= rng.choice([-1.0, 1.0], (T, 24, 1))                             # un-canonicalised signs, like the real clip
        if i % 3 == 2 and T > 6:
            q[2:5] = q[2]                                                    # identical consecutive frames
        tr = np.cumsum(rng.normal(0, 0.02, (T, 3)), 0) + np.array([0, 0, 0.9])
        clips[f"s{i}"] = {"root_trans_offset": tr, "pose_aa": rng.normal(0, 0.3, (T, 72)), "pose_quat_global": q,
                          "beta": np.zeros(16), "gender": "neutral", "fps": int(fps[i % len(fps)])}
    return RawClips.from_dict(clips)


@pytest.mark.gpu
def test_build_matches_oracle_at_every_tile_boundary(g):
    lengths = [2, 3, 9, 17, 18, 31, 32, 33, 41, 63, 64, 65, 96, 97, 150, 299, 300, 301, 700]
    raw = _synthetic_raw(lengths)
    lib = _lib(raw, max_length=300)
    n = len(lengths)
    lib.load_motions(skeleton_trees=[_skeleton(g)] * n, gender_betas=torch.zeros(n, 17), limb_weights=np.zeros((n, 10)), random_sample=False)
    T = _tables_np(lib)
    assert T["num_frames"].tolist() == [min(x, 300) for x in lengths]
    for i in range(n):
        c, nf, a = raw.clip(i), int(T["num_frames"][i]), int(T["length_starts"][i])
        o = co.build_clip(c["pose_quat_global"][:nf], c["root_trans_offset"][:nf], g["parents"], g["local_translation"], c["fps"])
        for k in BITS:
            assert_equal(_bits(T[k][a:a + nf]), _bits(o[k]), f"len {lengths[i]} {k} bits")
        for k in ("gvs", "gavs"):
            assert_close(T[k][a:a + nf], o[k], rtol=1e-6, atol=1e-7, what=f"len {lengths[i]} {k}")
        assert_close(T["dvs"][a:a + nf], o["dvs"], rtol=1e-5, atol=1e-6, what=f"len {lengths[i]} dvs")


@pytest.mark.gpu
def test_per_slot_skeletons_and_dedupe(g):
    """Different skeleton per slot (lt_clip_stride != 0); de-duplicated slots share rows and answer queries identically."""
    from puffer_phc_b200.skeleton import SkeletonTree
    raw = _synthetic_raw([40, 75, 33], seed=5, fps=(30,))
    lib = _lib(raw, max_length=60)
    sks = [SkeletonTree([f"b{j}" for j in range(24)], g["parents"].astype(np.int32), g["local_translation"] * s) for s in (1.0, 1.1, 0.9, 1.0, 1.1)]
    idx = torch.tensor([0, 1, 2, 1, 0])
    lib.load_motions(skeleton_trees=sks, gender_betas=torch.zeros(5, 17), limb_weights=np.zeros((5, 10)), sample_idxes=idx)
    T = _tables_np(lib)
    for s in range(5):
        c, nf, a = raw.clip(int(idx[s])), int(T["num_frames"][s]), int(T["length_starts"][s])
        o = co.build_clip(c["pose_quat_global"][:nf], c["root_trans_offset"][:nf], g["parents"], sks[s].local_translation.numpy(), c["fps"])
        assert_equal(_bits(T["gts"][a:a + nf]), _bits(o["gts"]), f"slot {s} gts bits")
    # dedupe: same skeleton everywhere, 64 slots over 3 clips
    n = 64
    idx = torch.arange(n) % 3
    sk = _skeleton(g)
    ids = torch.arange(n, device=DEV)
    times = torch.rand(n, device=DEV) * 1.0
    lib.load_motions(skeleton_trees=[sk] * n, gender_betas=torch.zeros(n, 17), limb_weights=np.zeros((n, 10)), sample_idxes=idx)
    full_rows = lib.gts.shape[0]
    a = lib.get_motion_state(ids, times)
    lib.load_motions(skeleton_trees=[sk] * n, gender_betas=torch.zeros(n, 17), limb_weights=np.zeros((n, 10)), sample_idxes=idx, dedupe=True)
    assert lib.gts.shape[0] == 40 + 60 + 33 and full_rows > 20 * lib.gts.shape[0]
    b = lib.get_motion_state(ids, times)
    for k in a:
        if k != "motion_aa":
            assert torch.equal(a[k], b[k]), k


@pytest.mark.gpu
def test_load_data_filters_and_sorts(g):
    from puffer_phc_b200.motion_file import RawClips
    raw = RawClips.from_dict({f"clip{i}": c for i, c in enumerate(_clips(g))})
    assert _lib(raw, min_length=40)._motion_data_keys.tolist() == ["clip0", "clip1"]                 # motion_lib.py:203-207
    assert _lib(raw, im_eval=True)._motion_data_keys.tolist() == ["clip0", "clip1", "clip3", "clip2"]  # :208-217 longest first
    lib = _lib(raw)
    assert lib._num_unique_motions == 4 and len(lib._motion_data_list) == 4
    with pytest.raises(ValueError):
        lib.load_motions(skeleton_trees=[SimpleNamespace(node_names=["a"] * 23)], gender_betas=torch.zeros(1, 17), limb_weights=np.zeros((1, 10)))


@pytest.mark.gpu
def test_full_size_library_properties(g):
    """BASELINE scale (11313 clips, ~2.5 M frames): size-independent properties of the table build, plus the oracle on a
    sample of clips spread over the library."""
    from puffer_phc_b200 import synth
    from puffer_phc_b200.motion_file import RawClips
    T = synth.make_motion_library(11313, seed=0, device=DEV)
    nf = T["num_frames"].cpu().numpy()
    fps = np.round(1.0 / T["motion_dt"].cpu().numpy()).astype(np.int32)
    raw = RawClips.from_device([f"c{i}" for i in range(len(nf))], nf, fps, T["gts"][:, 0].double().contiguous(),
                               T["motion_aa"].double().contiguous(), T["grs"].double().contiguous())
    del T
    lib = _lib(raw, max_length=300)
    n = len(nf)
    sk = _skeleton(g)
    lib.load_motions(skeleton_trees=[sk] * n, gender_betas=torch.zeros(n, 17), limb_weights=np.zeros((n, 10)), random_sample=False)
    d = raw.to_device(DEV)
    F = int(nf.sum())
    assert lib.gts.shape == (F, 24, 3) and lib.packed.shape == (F, 312)
    assert torch.equal(lib.length_starts.cpu(), torch.from_numpy(np.concatenate([[0], np.cumsum(nf)[:-1]])))
    # the root keeps float32(root translation), global rotations are the float32 cast of the input
    assert torch.equal(lib.gts[:, 0], d["root_trans"].float())
    assert torch.equal(lib.grs, d["pose_quat_global"].float())
    # local rotations are unit quaternions with w >= 0 (quat_normalize) except the root, which keeps the input
    assert float((lib.lrs[:, 1:].norm(dim=-1) - 1).abs().max()) < 1e-6 and float(lib.lrs[:, 1:, 3].min()) >= 0.0
    # bone lengths are preserved by the forward kinematics
    par = torch.from_numpy(g["parents"][1:]).to(DEV)
    bone = (lib.gts[:, 1:] - lib.gts[:, par]).norm(dim=-1)
    want = torch.from_numpy(np.linalg.norm(g["local_translation"][1:], axis=-1)).to(DEV)
    assert float((bone - want).abs().max()) < 1e-5
    # the last frame of every clip repeats the dof velocity of the one before it; the packed records mirror the tables
    last = torch.from_numpy(np.cumsum(nf) - 1).to(DEV)
    assert torch.equal(lib.dvs[last], lib.dvs[last - 1])
    for k, (lo, hi) in {"gts": (0, 72), "grs": (72, 168), "gvs": (168, 240), "gavs": (240, 312)}.items():
        assert torch.equal(lib.packed[:, lo:hi], getattr(lib, k).reshape(F, -1)), k
    assert all(bool(torch.isfinite(getattr(lib, k)).all()) for k in FIELDS)
    assert torch.equal(lib.grvs, lib.gvs[:, 0]) and torch.equal(lib.gravs, lib.gavs[:, 0])
    # oracle on every 400th clip
    starts = lib.length_starts.cpu().numpy()
    for i in range(0, n, 400):
        c, a, m = raw.clip(i), int(starts[i]), int(nf[i])
        o = co.build_clip(c["pose_quat_global"].cpu().numpy(), c["root_trans_offset"].cpu().numpy(), g["parents"], g["local_translation"], c["fps"])
        for k in BITS:
            assert_equal(_bits(getattr(lib, k)[a:a + m].cpu().numpy()), _bits(o[k]), f"clip {i} {k} bits")
        for k in ("gvs", "gavs"):
            assert_close(getattr(lib, k)[a:a + m].cpu().numpy(), o[k], rtol=1e-6, atol=1e-7, what=f"clip {i} {k}")
        assert_close(lib.dvs[a:a + m].cpu().numpy(), o["dvs"], rtol=1e-5, atol=1e-6, what=f"clip {i} dvs")


@pytest.mark.gpu
def test_dedupe_respects_per_slot_skeletons(g):
    """Slots that load the same clip with DIFFERENT skeleton objects must not share rows; with dedupe they still answer queries
    exactly like the un-deduplicated load."""
    from puffer_phc_b200.skeleton import SkeletonTree
    raw = _synthetic_raw([40, 52], seed=9, fps=(30,))
    lib = _lib(raw, max_length=60)
    mk = lambda s: SkeletonTree([f"b{j}" for j in range(24)], g["parents"].astype(np.int32), g["local_translation"] * s)   # noqa: E731
    a, b = mk(1.0), mk(1.2)
    sks = [a, b, a, b, a, a]
    idx = torch.tensor([0, 0, 0, 1, 1, 1])
    kw = dict(skeleton_trees=sks, gender_betas=torch.zeros(6, 17, device=DEV), limb_weights=torch.zeros(6, 10), sample_idxes=idx)
    ids, times = torch.arange(6, device=DEV), torch.linspace(0.0, 1.2, 6, device=DEV)
    lib.load_motions(**kw)
    want = lib.get_motion_state(ids, times)
    lib.load_motions(dedupe=True, **kw)
    assert lib.gts.shape[0] == 40 + 40 + 52 + 52          # (clip 0, a), (clip 0, b), (clip 1, b), (clip 1, a)
    got = lib.get_motion_state(ids, times)
    for k in want:
        if k != "motion_aa":
            assert torch.equal(want[k], got[k]), k
    assert not torch.equal(got["rg_pos"][0], got["rg_pos"][1])     # same clip and time, different skeleton
