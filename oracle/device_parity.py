"""TEST INFRASTRUCTURE -- the kernels against the reference's OWN functions (oracle/_ref, see ref_runner.py) executed on the box on
torch-CPU and on torch-CUDA, per output, with mismatch counts for integers / flags and worst margins for floats.

Used by tests/test_reference_devices.py (assertions) and profiles/tools/parity_report.py (the committed report).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import ref_runner as rr

RTOL, ATOL = 1e-5, 2e-6

# Layout of the 934-float observation row: (name, first column, width of one vector, vectors)
OBS_BLOCKS = (("root_h", 0, 1, 1), ("self_pos", 1, 3, 23), ("self_rot", 70, 6, 24), ("self_vel", 214, 3, 24), ("self_ang", 286, 3, 24),
              ("d_pos", 358, 3, 24), ("d_rot", 430, 6, 24), ("d_vel", 574, 3, 24), ("d_ang", 646, 3, 24), ("l_pos", 718, 3, 24),
              ("l_rot", 790, 6, 24))


def obs_floor(ref: np.ndarray, atol: float = ATOL) -> np.ndarray:
    """Absolute floor per observation element: atol * max(1, |v|_inf) with v the 3- or 6-vector the element belongs to.  A component
    of a rotated vector that sits near zero carries an absolute error of ~eps * |v| in ANY fp32 implementation (torch-CUDA and
    torch-CPU differ from each other in exactly this way), so the floor scales with the vector it came from -- never with other
    vectors of the row (metres, unit axes and rad/s do not mix)."""
    a = np.abs(np.asarray(ref, np.float64))
    floor = np.empty_like(a)
    for _, c0, w, n in OBS_BLOCKS:
        vmax = a[:, c0:c0 + w * n].reshape(a.shape[0], n, w).max(axis=-1, keepdims=True)
        floor[:, c0:c0 + w * n] = np.broadcast_to(atol * np.maximum(1.0, vmax), (a.shape[0], n, w)).reshape(a.shape[0], n * w)
    return floor


def err_over_tol(got, ref, floor=ATOL, rtol=RTOL):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.abs(got - ref) / (rtol * np.abs(ref) + floor)


_LIBS: Dict = {}


def ref_lib(T_dev: Dict[str, torch.Tensor], device):
    """The reference MotionLib over the given tables on ``device`` (cached: moving 5.5 GB of tables to the host takes seconds)."""
    key = (T_dev["gts"].data_ptr(), str(device))
    if key not in _LIBS:
        _LIBS.clear()                 # keep one library resident at a time per process
        _LIBS[key] = rr.lib_from_tables(T_dev, device)
    return _LIBS[key]


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def compare_step(T_dev: Dict[str, torch.Tensor], S_cpu: Dict[str, torch.Tensor], flavour: str, eval_mode: bool = False,
                 ours_lib=None, warm: int = 3, eval_distance: float = 0.5) -> Dict:
    """Run the reference's step on ``flavour`` ("cpu" | "cuda") and the fused kernel with ``ref_device=flavour``; return the report.
    ``T_dev``: tables on the CUDA device; ``S_cpu``: synthetic env state on the host."""
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    dev = T_dev["gts"].device
    rdev = dev if flavour == "cuda" else torch.device("cpu")
    lib_r = ref_lib(T_dev, rdev)
    S_r = {k: v.to(rdev) for k, v in S_cpu.items()}
    for _ in range(warm if flavour == "cuda" else 1):      # TorchScript's profiling executor specialises / fuses after two runs
        ref = rr.step(lib_r, S_r, eval_mode=eval_mode, with_blend=True, eval_distance=eval_distance)
    marg = rr.flag_margins(lib_r, S_r)
    if ours_lib is None:
        ours_lib = MotionLibSMPL.from_tables(T_dev, device=dev)
    N = S_cpu["progress"].shape[0]
    cfg = StepConfig(ref_device=flavour)
    if eval_mode:
        import dataclasses
        cfg = dataclasses.replace(cfg.eval_mode(), termination_distance=float(eval_distance))
    fs = FusedStep(ours_lib, N, cfg)
    Sg = {k: v.to(dev) for k, v in S_cpu.items()}
    o = fs(Sg["body_state"], Sg["progress"], Sg["start_time"], Sg["start_offset"], Sg["motion_ids"], Sg["global_offset"],
           Sg["dof_force"], Sg["dof_vel"])
    import puffer_phc_b200
    prev = puffer_phc_b200.set_reference_device(flavour)
    try:
        rep = {"flavour": flavour, "envs": int(N), "eval_mode": bool(eval_mode)}
        # frame indices / blend through the drop-in get_motion_state (debug outputs of _calc_frame_blend)
        for tag, step_add in (("t0", 0), ("t1", 1)):
            t = (Sg["progress"] + step_add) * rr.DT + Sg["start_time"] + Sg["start_offset"]
            ms = ours_lib.get_motion_state(Sg["motion_ids"], t, Sg["global_offset"], keys=("rg_pos",), debug=True)
            for k in ("idx0", "idx1"):
                rep[f"{tag}_{k}_mismatches"] = int((_np(ms[f"frame_{k}"]) != _np(ref[f"{tag}_{k}"])).sum())
            rep[f"{tag}_blend_bit_mismatches"] = int((_np(ms["blend"]).view(np.uint32) != _np(ref[f"{tag}_blend"]).view(np.uint32)).sum())
    finally:
        puffer_phc_b200.set_reference_device(prev)
    for k in ("reset", "terminated"):
        rep[f"{k}_mismatches"] = int((_np(o[k]).astype(bool) != _np(ref[k]).astype(bool)).sum())
        rep[f"{k}_set"] = int(_np(ref[k]).sum())
    dm, tm = _np(marg["dist_margin"]), _np(marg["time_margin"])
    rep["termination_margin_min_m"] = float(dm.min())                   # closest |distance - threshold| of any env (train variant)
    rep["termination_margin_lt_1e-6"] = int((dm < 1e-6).sum())
    rep["pass_time_margin_min_s"] = float(tm.min())
    rep["pass_time_margin_lt_1e-6"] = int((tm < 1e-6).sum())
    obs_ref = _np(ref["obs"])
    e_plain = err_over_tol(_np(o["obs"]), obs_ref)
    e_block = err_over_tol(_np(o["obs"]), obs_ref, obs_floor(obs_ref))
    rep["obs"] = {"plain_tol_worst": float(e_plain.max()), "plain_tol_n_over": int((e_plain > 1).sum()),
                  "vector_floor_worst": float(e_block.max()), "vector_floor_n_over": int((e_block > 1).sum()), "n": int(e_plain.size)}
    rep["obs_blocks_plain_worst"] = {name: float(e_plain[:, c0:c0 + w * n].max()) for name, c0, w, n in OBS_BLOCKS}
    for k in ("reward", "reward_raw"):
        e = err_over_tol(_np(o[k]), _np(ref[k]))
        rep[k] = {"plain_tol_worst": float(e.max()), "plain_tol_n_over": int((e > 1).sum()), "n": int(e.size)}
    return rep, {"ours": {k: _np(v) for k, v in o.items()}, "ref": {k: _np(v) for k, v in ref.items()}}
