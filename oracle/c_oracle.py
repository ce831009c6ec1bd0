"""TEST INFRASTRUCTURE -- ctypes binding of oracle/phc_oracle.c (the plain-C CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
All arguments are numpy arrays (C-contiguous, dtypes as documented); outputs are new numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libphc_oracle.so")
_lib = None

STATE_KEYS = (
    "root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "motion_aa",
    "rg_pos", "rb_rot", "body_vel", "body_ang_vel", "motion_bodies", "motion_limb_weights",
)
_STATE_SHAPES = {
    "root_pos": (3,), "root_rot": (4,), "dof_pos": (69,), "root_vel": (3,), "root_ang_vel": (3,),
    "dof_vel": (69,), "motion_aa": (72,), "rg_pos": (24, 3), "rb_rot": (24, 4), "body_vel": (24, 3),
    "body_ang_vel": (24, 3), "motion_bodies": (17,), "motion_limb_weights": (10,),
}
TABLE_KEYS = ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "motion_aa", "motion_len", "motion_dt",
              "num_frames", "length_starts", "motion_bodies", "limb_weights")


def build(force: bool = False) -> str:
    """Compile oracle/phc_oracle.c + loader_oracle.c -> oracle/_build/libphc_oracle.so (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("phc_oracle.c", "loader_oracle.c")]
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "CC=gcc"])
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def set_ref_device(kind) -> None:
    """Whose torch rounding the oracle's device-dependent reductions follow: "cpu" (default; the golden vectors) or "cuda"."""
    lib().phc_oracle_set_ref_device(1 if kind in ("cuda", 1) else 0)


class _Tables(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in TABLE_KEYS] + [("F", C.c_int64), ("M", C.c_int64)]


class _StateOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in STATE_KEYS] + [("idx0", C.c_void_p), ("idx1", C.c_void_p), ("blend", C.c_void_p)]


class _StepIn(C.Structure):
    _fields_ = [
        ("body_state", C.c_void_p), ("env_stride", C.c_int64),
        ("progress", C.c_void_p), ("start_time", C.c_void_p), ("start_offset", C.c_void_p), ("motion_ids", C.c_void_p),
        ("global_offset", C.c_void_p), ("dof_force", C.c_void_p), ("dof_vel", C.c_void_p),
        ("dt", C.c_float), ("k", C.c_float * 4), ("w", C.c_float * 4), ("power_coef", C.c_float),
        ("term_dist", C.c_void_p), ("reset_body_mask", C.c_uint32),
        ("enable_early_termination", C.c_int), ("use_mean", C.c_int), ("N", C.c_int64),
    ]


class _StepOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("obs", "reward", "reward_raw", "reset", "terminated", "ref_t", "ref_t1")]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    """Address of a numpy array as a c_void_p (a bare int would be truncated to 32 bits by ctypes)."""
    return C.c_void_p(None if a is None else a.ctypes.data)


class Tables:
    """Motion tables (A0 in SURVEY.md section 8a; motion_lib.py:396-420) held as fp32/int64 numpy arrays."""

    def __init__(self, **kw):
        self.arr = {}
        for k in TABLE_KEYS:
            v = kw[k]
            self.arr[k] = _i64(v) if k in ("num_frames", "length_starts") else _f32(v)
        self.F = int(self.arr["gts"].shape[0])
        self.M = int(self.arr["motion_len"].shape[0])
        self.c = _Tables(*[self.arr[k].ctypes.data for k in TABLE_KEYS], self.F, self.M)

    @classmethod
    def from_npz(cls, path):
        z = np.load(path)
        return cls(**{k: z[k] for k in TABLE_KEYS})


def frame_blend(time, length, num_frames, dt):
    time, length, dt, nf = _f32(time), _f32(length), _f32(dt), _i64(num_frames)
    n = time.shape[0]
    i0, i1, bl = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float32)
    lib().phc_oracle_frame_blend(_p(time), _p(length), _p(nf), _p(dt), C.c_int64(n), _p(i0), _p(i1), _p(bl))
    return i0, i1, bl


def motion_state(tables: Tables, motion_ids, motion_times, offset=None, debug=False):
    ids, times = _i64(motion_ids), _f32(motion_times)
    B = ids.shape[0]
    off = None if offset is None else _f32(offset)
    out = {k: np.empty((B,) + _STATE_SHAPES[k], np.float32) for k in STATE_KEYS}
    i0, i1, bl = np.empty(B, np.int64), np.empty(B, np.int64), np.empty(B, np.float32)
    so = _StateOut(*[out[k].ctypes.data for k in STATE_KEYS], _p(i0), _p(i1), _p(bl))
    lib().phc_oracle_motion_state(C.byref(tables.c), _p(ids), _p(times), _p(off), C.c_int64(B), C.byref(so))
    if debug:
        return out, (i0, i1, bl)
    return out


def sample_time_interval(phase, motion_len, div_mode=0):
    phase, motion_len = _f32(phase), _f32(motion_len)
    out = np.empty_like(phase)
    lib().phc_oracle_sample_time_interval(_p(phase), _p(motion_len), C.c_int64(phase.shape[0]), C.c_int(div_mode), _p(out))
    return out


def imitation_obs_v6(root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel,
                     ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel, time_steps=1, upright=True):
    assert time_steps == 1
    a = [_f32(x) for x in (root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel,
                           ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel)]
    N, J = a[2].shape[0], a[2].shape[1]
    obs = np.empty((N, J * 24), np.float32)
    lib().phc_oracle_imitation_obs_v6(*[_p(x) for x in a], C.c_int64(N), C.c_int(J), C.c_int(int(upright)), _p(obs))
    return obs


def self_obs(body_pos, body_rot, body_vel, body_ang_vel, local_root_obs=True, root_height_obs=True, upright=True):
    a = [_f32(x) for x in (body_pos, body_rot, body_vel, body_ang_vel)]
    N, J = a[0].shape[0], a[0].shape[1]
    W = (1 if root_height_obs else 0) + 3 * (J - 1) + 12 * J
    obs = np.empty((N, W), np.float32)
    lib().phc_oracle_self_obs(*[_p(x) for x in a], C.c_int64(N), C.c_int(J), C.c_int(int(local_root_obs)),
                              C.c_int(int(root_height_obs)), C.c_int(int(upright)), _p(obs))
    return obs


def imitation_reward(body_pos, body_rot, body_vel, body_ang_vel, ref_body_pos, ref_body_rot, ref_body_vel,
                     ref_body_ang_vel, k, w):
    a = [_f32(x) for x in (body_pos, body_rot, body_vel, body_ang_vel, ref_body_pos, ref_body_rot, ref_body_vel,
                           ref_body_ang_vel)]
    N, J = a[0].shape[0], a[0].shape[1]
    k, w = _f32(k), _f32(w)
    rew, raw = np.empty(N, np.float32), np.empty((N, 4), np.float32)
    lib().phc_oracle_imitation_reward(*[_p(x) for x in a], C.c_int64(N), C.c_int(J), _p(k), _p(w), _p(rew), _p(raw))
    return rew, raw


def im_reset(progress, body_pos, ref_body_pos, pass_time, enable_early_termination, termination_distance, use_mean):
    prog = np.ascontiguousarray(progress, dtype=np.int16)
    bp, rp = _f32(body_pos), _f32(ref_body_pos)
    pt = np.ascontiguousarray(pass_time, dtype=np.uint8)
    td = _f32(termination_distance)
    N, J = bp.shape[0], bp.shape[1]
    reset, term = np.empty(N, np.uint8), np.empty(N, np.uint8)
    lib().phc_oracle_im_reset(_p(prog), _p(bp), _p(rp), _p(pt), C.c_int(int(enable_early_termination)), _p(td),
                              C.c_int(int(use_mean)), C.c_int64(N), C.c_int(J), _p(reset), _p(term))
    return reset.astype(bool), term.astype(bool)


def step(tables: Tables, body_state, progress, start_time, start_offset, motion_ids, global_offset, dt,
         k, w, term_dist, reset_body_mask=0xFFFFFF, enable_early_termination=True, use_mean=False,
         dof_force=None, dof_vel=None, power_coef=0.0005, want_ref=False):
    """Full post-physics step (humanoid_phc.py:136-149).  body_state: [N, S>=312] AoS records."""
    bs = _f32(body_state)
    N = bs.shape[0]
    bs2 = bs.reshape(N, -1)
    prog = np.ascontiguousarray(progress, dtype=np.int16)
    st, so, ids, go = _f32(start_time), _f32(start_offset), _i64(motion_ids), _f32(global_offset)
    td = _f32(term_dist)
    df = None if dof_force is None else _f32(dof_force)
    dv = None if dof_vel is None else _f32(dof_vel)
    rw = 5 if df is not None else 4
    obs, rew, raw = np.empty((N, 934), np.float32), np.empty(N, np.float32), np.empty((N, rw), np.float32)
    reset, term = np.empty(N, np.uint8), np.empty(N, np.uint8)
    ref_t = np.empty((N, 312), np.float32) if want_ref else None
    ref_t1 = np.empty((N, 312), np.float32) if want_ref else None
    sin = _StepIn(_p(bs2), bs2.shape[1], _p(prog), _p(st), _p(so), _p(ids), _p(go), _p(df), _p(dv),
                  float(np.float32(dt)), (C.c_float * 4)(*[float(x) for x in k]), (C.c_float * 4)(*[float(x) for x in w]),
                  float(np.float32(power_coef)), _p(td), reset_body_mask, int(enable_early_termination), int(use_mean), N)
    sout = _StepOut(_p(obs), _p(rew), _p(raw), _p(reset), _p(term), _p(ref_t), _p(ref_t1))
    lib().phc_oracle_step(C.byref(tables.c), C.byref(sin), C.byref(sout))
    res = dict(obs=obs, reward=rew, reward_raw=raw, reset=reset.astype(bool), terminated=term.astype(bool))
    if want_ref:
        res["ref_t"], res["ref_t1"] = ref_t, ref_t1
    return res


def rms_forward(x, mean, var, eps=1e-5, clip=10.0):
    x, mean, var = _f32(x), _f32(mean).reshape(-1), _f32(var).reshape(-1)
    y = np.empty_like(x)
    lib().phc_oracle_rms_forward(_p(x), _p(mean), _p(var), C.c_float(eps), C.c_float(clip),
                                 C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(y))
    return y


def rms_update(x, running_mean, running_var, count):
    """Returns new (running_mean [1,C], running_var [1,C], count [1]); inputs are not modified."""
    x = _f32(x)
    m, v, c = _f32(running_mean).reshape(-1).copy(), _f32(running_var).reshape(-1).copy(), _f32(count).reshape(-1).copy()
    lib().phc_oracle_rms_update(_p(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _p(m), _p(v), _p(c))
    return m.reshape(1, -1), v.reshape(1, -1), c


def gae(dones, values, rewards, gamma, gae_lambda):
    d, v, r = _f32(dones), _f32(values), _f32(rewards)
    adv = np.empty_like(r)
    lib().phc_oracle_gae(_p(d), _p(v), _p(r), C.c_int64(r.shape[0]), C.c_float(gamma), C.c_float(gae_lambda), _p(adv))
    return adv


def amp_obs(root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos, dof_subset=None,
            local_root_obs=True, root_height_obs=True, upright=True):
    a = [_f32(x) for x in (root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos)]
    N, K = a[0].shape[0], a[6].shape[1]
    sub = None if dof_subset is None else _i64(dof_subset)
    nj = 23 if sub is None else sub.shape[0] // 3
    W = (1 if root_height_obs else 0) + 12 + 9 * nj + 3 * K
    obs = np.empty((N, W), np.float32)
    lib().phc_oracle_amp_obs(*[_p(x) for x in a], _p(sub), C.c_int(nj), C.c_int(K), C.c_int(int(local_root_obs)),
                             C.c_int(int(root_height_obs)), C.c_int(int(upright)), C.c_int64(N), _p(obs))
    return obs


def build_clip(pose_quat_global, root_trans, parents, local_translation, fps):
    """Motion table build for one (already cropped) clip -- oracle/loader_oracle.c.  Inputs float64 as in the pkl format;
    returns the float32 per-clip table slices dict(gts, grs, lrs, gvs, gavs, dvs)."""
    q = np.ascontiguousarray(pose_quat_global, dtype=np.float64)
    tr = np.ascontiguousarray(root_trans, dtype=np.float64)
    T, J = q.shape[0], q.shape[1]
    par, lt = _i64(parents), _f32(local_translation)
    out = dict(gts=np.empty((T, J, 3), np.float32), grs=np.empty((T, J, 4), np.float32), lrs=np.empty((T, J, 4), np.float32),
               gvs=np.empty((T, J, 3), np.float32), gavs=np.empty((T, J, 3), np.float32), dvs=np.empty((T, J - 1, 3), np.float32))
    rc = lib().phc_oracle_build_clip(_p(q), _p(tr), C.c_int64(T), C.c_int(J), _p(par), _p(lt), C.c_int(int(fps)),
                                     *[_p(out[k]) for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs")])
    if rc != 0:
        raise ValueError(f"phc_oracle_build_clip failed ({rc}): clips need at least 2 frames")
    return out


def mpjpe(body_pos, ref_body_pos):
    bp, rp = _f32(body_pos), _f32(ref_body_pos)
    out = np.empty(bp.shape[0], np.float32)
    lib().phc_oracle_mpjpe(_p(bp), _p(rp), C.c_int64(bp.shape[0]), C.c_int(bp.shape[1]), _p(out))
    return out
