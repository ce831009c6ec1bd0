"""ctypes binding of libphc_b200.so (include/phc_b200.h).  No torch types cross the ABI: tensors are passed
as ``data_ptr()`` integers plus sizes/strides, the stream as ``torch.cuda.current_stream().cuda_stream``.

There is deliberately NO fallback: if the CUDA library cannot be loaded every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

c_f32p, c_i64p, c_u8p, c_i16p, c_f64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p

EXPORTS = (
    "phc_version", "phc_last_error", "phc_pack_frames", "phc_build_pair_aux", "phc_motion_state", "phc_reset_ref_state", "phc_sample_time_interval",
    "phc_imitation_obs_v6", "phc_self_obs_smpl_max", "phc_amp_obs_smpl", "phc_amp_obs_hist_step", "phc_imitation_reward", "phc_im_reset",
    "phc_step_num_partials", "phc_step_fused", "phc_rms_forward", "phc_rms_scratch_doubles", "phc_rms_moments",
    "phc_rms_reduce_partials", "phc_rms_finalize", "phc_auto_reset_num_partials", "phc_auto_reset_scratch_bytes", "phc_auto_reset",
    "phc_stats_reduce", "phc_stats_comm_bytes", "phc_stats_allreduce_finalize", "phc_gae", "phc_build_motion_tables", "phc_build_motion_aa", "phc_cast_f64_f32", "phc_mpjpe", "phc_frame_blend",
    "phc_rollout_store", "phc_rollout_scratch_bytes", "phc_rollout_sort",
)

VERSION = 121
NUM_METRICS = 16
METRIC_NAMES = ("steps", "reward", "r_pos", "r_rot", "r_vel", "r_ang_vel", "r_power", "resets", "terminations", "truncations",
                "episode_return", "episode_length", "episodes")
REF_CPU, REF_CUDA = 0, 1
PHC_OK, PHC_EINVAL, PHC_EALIGN, PHC_ESHAPE, PHC_EUNSUPPORTED = 0, -1, -2, -3, -4


class View(C.Structure):
    """phc_view: [N, J, C] fp32 tensor with unit innermost stride (strides in floats)."""
    _fields_ = [("ptr", C.c_void_p), ("stride_env", C.c_int64), ("stride_body", C.c_int64)]


TABLE_FIELDS = ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "motion_aa", "motion_len", "motion_dt", "num_frames",
                "length_starts", "motion_bodies", "limb_weights", "packed")


class MotionTables(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in TABLE_FIELDS] + [("pair_aux", C.c_void_p), ("pair_flags", C.c_void_p), ("pair_device", C.c_int),
                                                          ("F", C.c_int64), ("M", C.c_int64)]


STATE_FIELDS = ("root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "motion_aa", "rg_pos", "rb_rot",
                "body_vel", "body_ang_vel", "motion_bodies", "motion_limb_weights", "frame_idx0", "frame_idx1", "blend")


class MotionStateOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in STATE_FIELDS]


class StepIn(C.Structure):
    _fields_ = [
        ("body_state", C.c_void_p), ("env_stride", C.c_int64), ("progress", C.c_void_p), ("start_time", C.c_void_p),
        ("start_offset", C.c_void_p), ("motion_ids", C.c_void_p), ("global_offset", C.c_void_p), ("dof_force", C.c_void_p),
        ("dof_vel", C.c_void_p), ("term_dist", C.c_void_p), ("rms_mean", C.c_void_p), ("rms_var", C.c_void_p), ("N", C.c_int64),
    ]


class StepCfg(C.Structure):
    _fields_ = [
        ("dt", C.c_float), ("k", C.c_float * 4), ("w", C.c_float * 4), ("power_coef", C.c_float),
        ("reset_body_mask", C.c_uint32), ("enable_early_termination", C.c_int), ("use_mean", C.c_int),
        ("rms_eps", C.c_float), ("rms_clip", C.c_float), ("ref_device", C.c_int),
    ]


class StepOut(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("obs_stride", C.c_int64), ("obs_norm", C.c_void_p), ("reward", C.c_void_p),
        ("reward_raw", C.c_void_p), ("raw_stride", C.c_int64), ("reset", C.c_void_p), ("terminated", C.c_void_p),
        ("moment_partials", C.c_void_p), ("accumulate_partials", C.c_int), ("ref_state_t", C.c_void_p), ("ref_state_t1", C.c_void_p),
        ("metric_partials", C.c_void_p),
    ]


class ResetEnv(C.Structure):
    _fields_ = [("body_state", C.c_void_p), ("env_stride", C.c_int64), ("root_states", C.c_void_p), ("dof_pos", C.c_void_p),
                ("dof_vel", C.c_void_p), ("progress", C.c_void_p), ("start_time", C.c_void_p), ("start_offset", C.c_void_p),
                ("global_offset", C.c_void_p), ("motion_ids", C.c_void_p), ("reset", C.c_void_p), ("terminated", C.c_void_p),
                ("obs", C.c_void_p), ("obs_stride", C.c_int64), ("obs_norm", C.c_void_p), ("rms_mean", C.c_void_p), ("rms_var", C.c_void_p)]


class ResetBook(C.Structure):
    _fields_ = [("rewards", C.c_void_p), ("reward_raw", C.c_void_p), ("raw_stride", C.c_int64), ("raw_dim", C.c_int),
                ("terminals", C.c_void_p), ("truncations", C.c_void_p), ("masks", C.c_void_p), ("episode_returns", C.c_void_p),
                ("episode_lengths", C.c_void_p), ("metrics", C.c_void_p), ("step_metrics", C.c_int)]


class ResetCfg(C.Structure):
    _fields_ = [("dt", C.c_float), ("state_init", C.c_int), ("flag_test", C.c_int), ("ref_device", C.c_int), ("rms_eps", C.c_float),
                ("rms_clip", C.c_float)]


class StatsComm(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("peer_bufs", C.c_void_p * 32), ("epoch", C.c_uint64), ("ticket", C.c_void_p)]


class BuildIn(C.Structure):
    """phc_build_in (raw clips -> tables, row f4)."""
    _fields_ = [(k, C.c_void_p) for k in ("pose_quat_global", "root_trans", "in_start", "num_frames", "out_start", "fps",
                                          "tile_prefix", "parents", "local_translation")] + \
               [("lt_clip_stride", C.c_int64), ("heading", C.c_void_p), ("M", C.c_int64), ("n_tiles", C.c_int64), ("J", C.c_int)]


class BuildOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "packed")]


BUILD_TILE = 32     # PHC_BUILD_TILE


def _declare(lib):
    P, I64, I, F = C.c_void_p, C.c_int64, C.c_int, C.c_float
    lib.phc_version.argtypes, lib.phc_version.restype = [], I
    lib.phc_last_error.argtypes, lib.phc_last_error.restype = [], C.c_char_p
    lib.phc_pack_frames.argtypes = [C.POINTER(MotionTables), P, P]
    lib.phc_build_pair_aux.argtypes = [C.POINTER(MotionTables), I, P, P, P]
    lib.phc_motion_state.argtypes = [C.POINTER(MotionTables), P, P, P, I64, C.POINTER(MotionStateOut), I, P]
    lib.phc_reset_ref_state.argtypes = [C.POINTER(MotionTables), P, P, P, P, I64, P, P, P, P, I64, I, P]
    lib.phc_sample_time_interval.argtypes = [P, P, I64, I, P, P]
    lib.phc_imitation_obs_v6.argtypes = [View] * 10 + [I64, I, I, I, P, I64, P]
    lib.phc_self_obs_smpl_max.argtypes = [View] * 4 + [I64, I, I, I, I, P, I64, P]
    lib.phc_amp_obs_smpl.argtypes = [P] * 8 + [I, I, I, I, I, I64, P, I64, I, P]
    lib.phc_amp_obs_hist_step.argtypes = [P] * 8 + [I, I, I, I, I, I64, P, I, I, I, P]
    lib.phc_imitation_reward.argtypes = [View] * 8 + [I64, I, C.POINTER(F), C.POINTER(F), P, P, I64, P]
    lib.phc_im_reset.argtypes = [P, View, View, P, I, P, I, I64, I, P, P, I, P]
    lib.phc_step_num_partials.argtypes, lib.phc_step_num_partials.restype = [], I
    lib.phc_step_fused.argtypes = [C.POINTER(MotionTables), C.POINTER(StepIn), C.POINTER(StepCfg), C.POINTER(StepOut), P]
    lib.phc_rms_forward.argtypes = [P, I64, P, P, F, F, I64, I, P, I64, P]
    lib.phc_rms_scratch_doubles.argtypes, lib.phc_rms_scratch_doubles.restype = [I], I64
    lib.phc_rms_moments.argtypes = [P, I64, I64, I, P, P, P]
    lib.phc_rms_reduce_partials.argtypes = [P, I, I64, I, P, P]
    lib.phc_rms_finalize.argtypes = [P, I, P, P, P, P]
    lib.phc_gae.argtypes = [P, P, P, I64, F, F, P, I, P]
    lib.phc_auto_reset_num_partials.argtypes, lib.phc_auto_reset_num_partials.restype = [], I
    lib.phc_auto_reset_scratch_bytes.argtypes, lib.phc_auto_reset_scratch_bytes.restype = [I64], I64
    lib.phc_auto_reset.argtypes = [C.POINTER(MotionTables), C.POINTER(ResetEnv), C.POINTER(ResetBook), C.POINTER(ResetCfg), P, I64, P, P, P,
                                   P, P, P]
    lib.phc_stats_reduce.argtypes = [P, I, I, I64, P, P, I, P, I, P]
    lib.phc_stats_comm_bytes.argtypes, lib.phc_stats_comm_bytes.restype = [I, I], I64
    lib.phc_stats_allreduce_finalize.argtypes = [P, I, I, I64, P, P, I, P, C.POINTER(StatsComm), P, P, P, P]
    lib.phc_build_motion_tables.argtypes = [C.POINTER(BuildIn), C.POINTER(BuildOut), P]
    lib.phc_build_motion_aa.argtypes = [P, I, P, P, I64, I64, P, P, P, P, P]
    lib.phc_cast_f64_f32.argtypes = [P, I64, P, P]
    lib.phc_mpjpe.argtypes = [View, View, I64, I, P, I, P]
    lib.phc_frame_blend.argtypes = [P, P, P, P, I64, P, P, P, P]
    lib.phc_rollout_store.argtypes = [P, P, P, P, I, P, I64, P, P, P, P, P, P, P, P, P, P]
    lib.phc_rollout_scratch_bytes.argtypes, lib.phc_rollout_scratch_bytes.restype = [I64, I], I64
    lib.phc_rollout_sort.argtypes = [P, P, P, P, P, P, P, I64, I, I64, P, P, P, P, P, P, P, P]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("phc_version", "phc_step_num_partials", "phc_auto_reset_num_partials"):
            fn.restype = I


# ---- reference device flavour (include/phc_b200.h: PHC_REF_DEVICE_*) ---------------------------------------------------
# The drop-in functions keep the reference's signatures, so the flavour they reproduce is a package-level setting.  Default: CUDA,
# the device the reference runs this path on; the committed golden vectors were made with torch-CPU, tests select "cpu" for those.
_ref_device = REF_CUDA


def set_reference_device(kind) -> int:
    """Select whose rounding flag-deciding reductions reproduce: ``"cuda"`` (default) or ``"cpu"``.  Returns the previous value."""
    global _ref_device
    prev = _ref_device
    if kind in ("cpu", REF_CPU):
        _ref_device = REF_CPU
    elif kind in ("cuda", REF_CUDA):
        _ref_device = REF_CUDA
    else:
        raise ValueError(f"reference device must be 'cpu' or 'cuda', got {kind!r}")
    return prev


def ref_device(override=None) -> int:
    if override is None:
        return _ref_device
    if override in ("cpu", REF_CPU):
        return REF_CPU
    if override in ("cuda", REF_CUDA):
        return REF_CUDA
    raise ValueError(f"reference device must be 'cpu' or 'cuda', got {override!r}")


def library_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (building if necessary) libphc_b200.so; raises RuntimeError if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = _build.LIB_PATH
            alt = os.environ.get("PHC_B200_LIB")          # tuning builds (profiles/tools/ab_variants.py): an explicit, prebuilt library
            if alt:
                if not os.path.isfile(alt):
                    raise RuntimeError(f"PHC_B200_LIB={alt} does not exist")
                path = alt
            elif not os.path.isfile(path) or _build.needs_build():
                try:
                    _build.build()
                except Exception as exc:  # no nvcc and no prebuilt library
                    if not os.path.isfile(path):
                        raise RuntimeError(
                            f"libphc_b200.so is missing and could not be built ({exc}); the CUDA kernels are the only "
                            "implementation of this package -- there is no CPU fallback") from exc
            lib = C.CDLL(path)
            _declare(lib)
            if lib.phc_version() != VERSION:
                raise RuntimeError(f"libphc_b200.so version {lib.phc_version()} does not match the Python package")
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().phc_last_error().decode("utf-8", "replace")
        kind = {PHC_EINVAL: "invalid argument", PHC_EALIGN: "alignment", PHC_ESHAPE: "shape",
                PHC_EUNSUPPORTED: "unsupported"}.get(rc, f"CUDA error {rc}")
        if rc == PHC_EUNSUPPORTED:
            raise NotImplementedError(f"{what or 'libphc_b200'}: {msg}")
        if rc < 0:
            raise ValueError(f"{what or 'libphc_b200'} ({kind}): {msg}")
        raise RuntimeError(f"{what or 'libphc_b200'} ({kind}): {msg}")


def require_cuda(*tensors):
    """The product path is CUDA-only: fail loudly rather than compute anything on the host."""
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("puffer_phc_b200: expected CUDA tensors -- the kernels are the only implementation "
                               f"(got a tensor on {t.device})")


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def on_device(dev):
    """Context that makes ``dev`` the current CUDA device for the launch -- a no-op object when it already is (the usual case:
    ``torch.cuda.device(...)`` alone costs several microseconds per call, more than a 4096-env kernel runs)."""
    import torch
    if not isinstance(dev, torch.device):
        dev = torch.device(dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return _NULL if idx == torch.cuda.current_device() else torch.cuda.device(idx)


def stream_ptr():
    """The current CUDA stream of the current device as a raw pointer (torch's C accessor: ~0.3 us; building a torch.cuda.Stream
    object for every call costs ~3.5 us, as much as the launch itself)."""
    import torch
    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(None if t is None else t.data_ptr())


def view3(t) -> View:
    """[N, J, C] (or [N, C]) fp32 tensor with unit inner stride -> phc_view; other layouts must be made contiguous first."""
    if t.dim() == 2:
        return View(t.data_ptr(), t.stride(0), 0)
    return View(t.data_ptr(), t.stride(0), t.stride(1))


def as_view_tensor(t):
    """Return a tensor that phc_view can describe without copying when possible (fp32, unit inner stride)."""
    import torch
    if t.dtype != torch.float32:
        t = t.float()
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        t = t.contiguous()
    return t
