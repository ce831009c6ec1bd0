"""Query side of the motion library: drop-in for the hot-path methods of the reference's
``MotionLibBase`` / ``MotionLibSMPL`` (reference puffer_phc/motion_lib.py:510-673).

Same method names, argument meaning, returned dict keys/shapes and attribute names the reference's
callers read (``_motion_lengths``, ``_num_unique_motions``, ``_curr_motion_ids``, ``_motion_data_keys``,
``_termination_history``).  The load side (pkl -> FK -> tables, motion_lib.py:180-429)
is out of scope for this round: tables are handed in ready-made (``from_tables``), exactly in the layout
``load_motions`` produces.  All arithmetic runs in ``libphc_b200.so`` (csrc/motion_state.cu); random
numbers stay on torch's generator so sampled ids / phases are bit-identical to the reference's.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _ffi

_TABLE_ATTR = {  # ABI field -> reference attribute name
    "gts": "gts", "grs": "grs", "lrs": "lrs", "gvs": "gvs", "gavs": "gavs", "dvs": "dvs", "motion_aa": "_motion_aa",
    "motion_len": "_motion_lengths", "motion_dt": "_motion_dt", "num_frames": "_motion_num_frames",
    "length_starts": "length_starts", "motion_bodies": "_motion_bodies", "limb_weights": "_motion_limb_weights",
}
_STATE_SHAPES = {
    "root_pos": (3,), "root_rot": (4,), "dof_pos": (69,), "root_vel": (3,), "root_ang_vel": (3,), "dof_vel": (69,),
    "motion_aa": (72,), "rg_pos": (24, 3), "rb_rot": (24, 4), "body_vel": (24, 3), "body_ang_vel": (24, 3),
    "motion_bodies": (17,), "motion_limb_weights": (10,),
}
STATE_KEYS = tuple(_STATE_SHAPES)


class MotionLibBase:
    """Device-resident motion tables + the reference's query API."""

    def __init__(self, tables: Dict[str, torch.Tensor], device=None, motion_data_keys=None, sim_fps: float = 30.0,
                 pack: bool = True):
        dev = torch.device(device) if device is not None else tables["gts"].device
        if dev.type != "cuda":
            raise RuntimeError("puffer_phc_b200.MotionLib: tables must live on a CUDA device (no CPU implementation)")
        self._device = dev
        self._sim_fps = sim_fps                       # motion_lib.py:183
        for field, attr in _TABLE_ATTR.items():
            t = tables[field]
            t = t.to(dev, dtype=torch.int64 if field in ("num_frames", "length_starts") else torch.float32).contiguous()
            setattr(self, attr, t)
        fps = tables.get("motion_fps")
        self._motion_fps = (1.0 / self._motion_dt) if fps is None else fps.to(dev, torch.float32)
        self._num_motions = int(self._motion_lengths.shape[0])
        self._num_unique_motions = self._num_motions
        self.num_bodies = 24
        self.num_joints = 24
        self.motion_ids = torch.arange(self._num_motions, dtype=torch.long, device=dev)           # :420
        self._curr_motion_ids = self.motion_ids.clone()
        self._motion_data_keys = motion_data_keys
        # sampling state (setup_constants, motion_lib.py:237-245)
        self._termination_history = torch.zeros(self._num_unique_motions, device=dev)
        self._success_rate = torch.zeros(self._num_unique_motions, device=dev)
        self._sampling_history = torch.zeros(self._num_unique_motions, device=dev)
        self._sampling_prob = torch.ones(self._num_unique_motions, device=dev) / self._num_unique_motions
        self._sampling_batch_prob = self._sampling_prob[self._curr_motion_ids] / self._sampling_prob[self._curr_motion_ids].sum()
        self.packed = None
        self._lib = _ffi.load()
        self._ctables = self._make_ctables()
        if pack:
            self.pack()

    # ------------------------------------------------------------------------------------------------
    @classmethod
    def from_tables(cls, tables: Dict[str, torch.Tensor], device=None, **kw) -> "MotionLibBase":
        return cls(tables, device=device, **kw)

    def _make_ctables(self) -> _ffi.MotionTables:
        vals = [getattr(self, _TABLE_ATTR[f]).data_ptr() for f in _ffi.TABLE_FIELDS[:-1]]
        packed = None if self.packed is None else self.packed.data_ptr()
        return _ffi.MotionTables(*vals, packed, int(self.gts.shape[0]), self._num_motions)

    def pack(self) -> torch.Tensor:
        """Build the B200 frame layout: one contiguous 1248-byte record (gts|grs|gvs|gavs) per frame."""
        F = int(self.gts.shape[0])
        packed = torch.empty((F, 312), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            _ffi.check(self._lib.phc_pack_frames(C.byref(self._ctables), _ffi.ptr(packed), _ffi.stream_ptr()), "phc_pack_frames")
        self.packed = packed
        self._ctables = self._make_ctables()
        return packed

    @property
    def ctables(self) -> _ffi.MotionTables:
        return self._ctables

    # ---- bookkeeping the reference exposes -----------------------------------------------------------
    def num_motions(self):
        return self._num_motions

    def get_total_length(self):
        return sum(self._motion_lengths)

    def get_motion_length(self, motion_ids=None):          # motion_lib.py:537-541
        return self._motion_lengths if motion_ids is None else self._motion_lengths[motion_ids]

    def get_motion_num_steps(self, motion_ids=None):       # motion_lib.py:543-547 (the ids form is broken upstream)
        if motion_ids is None:
            return (self._motion_num_frames * self._sim_fps / self._motion_fps).ceil().int()
        return (self._motion_num_frames[motion_ids] * self._sim_fps / self._motion_fps[motion_ids]).ceil().int()

    # ---- sampling: RNG stays torch's, arithmetic is ours -------------------------------------------
    def sample_motions(self, n):                           # motion_lib.py:510-513
        return torch.multinomial(self._sampling_batch_prob, num_samples=n, replacement=True).to(self._device)

    def sample_time(self, motion_ids, truncate_time=None):  # motion_lib.py:515-524
        phase = torch.rand(motion_ids.shape, device=self._device)
        motion_len = self._motion_lengths[motion_ids]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len -= truncate_time
        return phase * motion_len

    def sample_time_interval(self, motion_ids, truncate_time=None, cpu_division: bool = False):
        """motion_lib.py:526-535.  ``cpu_division`` selects the reference's CPU rounding (true division);
        the default reproduces what the reference computes when it runs on CUDA (scalar reciprocal multiply)."""
        phase = torch.rand(motion_ids.shape, device=self._device)
        motion_len = self._motion_lengths[motion_ids]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len -= truncate_time
        return self.time_interval_from_phase(phase, motion_len, cpu_division)

    def time_interval_from_phase(self, phase, motion_len, cpu_division: bool = False):
        _ffi.require_cuda(phase, motion_len)
        phase, motion_len = phase.contiguous().float(), motion_len.contiguous().float()
        out = torch.empty_like(phase)
        with torch.cuda.device(self._device):
            _ffi.check(self._lib.phc_sample_time_interval(_ffi.ptr(phase), _ffi.ptr(motion_len), phase.numel(),
                                                          0 if cpu_division else 1, _ffi.ptr(out), _ffi.stream_ptr()),
                       "phc_sample_time_interval")
        return out

    # ---- the hot query -----------------------------------------------------------------------------
    def get_motion_state(self, motion_ids, motion_times, offset=None, keys=None, debug: bool = False):
        """motion_lib.py:549-626.  Returns the reference's 13-key dict (``keys`` optionally restricts the
        outputs that are computed; ``debug`` adds ``frame_idx0/frame_idx1/blend`` from _calc_frame_blend)."""
        _ffi.require_cuda(motion_ids, motion_times, offset)
        ids = motion_ids.to(torch.int64).contiguous()
        times = motion_times.to(torch.float32).contiguous()
        off = None if offset is None else offset.to(torch.float32).contiguous()
        B = ids.shape[0]
        want = STATE_KEYS if keys is None else tuple(keys)
        out = {k: torch.empty((B,) + _STATE_SHAPES[k], dtype=torch.float32, device=self._device) for k in want}
        dbg = {}
        if debug:
            dbg = {"frame_idx0": torch.empty(B, dtype=torch.int64, device=self._device),
                   "frame_idx1": torch.empty(B, dtype=torch.int64, device=self._device),
                   "blend": torch.empty(B, dtype=torch.float32, device=self._device)}
        so = _ffi.MotionStateOut(*[(out[k].data_ptr() if k in out else None) for k in _ffi.STATE_FIELDS[:13]],
                                 *[(dbg[k].data_ptr() if k in dbg else None) for k in _ffi.STATE_FIELDS[13:]])
        with torch.cuda.device(self._device):
            _ffi.check(self._lib.phc_motion_state(C.byref(self._ctables), _ffi.ptr(ids), _ffi.ptr(times), _ffi.ptr(off), B,
                                                  C.byref(so), _ffi.stream_ptr()), "phc_motion_state")
        out.update(dbg)
        return out

    def get_root_pos_smpl(self, motion_ids, motion_times):   # motion_lib.py:628-653
        return self.get_motion_state(motion_ids, motion_times, offset=None, keys=("root_pos",))

    def _calc_frame_blend(self, time, len, num_frames, dt):   # noqa: A002  (reference signature, motion_lib.py:655-665)
        """Reference-signature helper; evaluated by the same kernel through a one-motion-per-row table view."""
        raise NotImplementedError("use get_motion_state(..., debug=True) to obtain frame_idx0/frame_idx1/blend")


class MotionLibSMPL(MotionLibBase):
    """Name the reference's callers import (motion_lib.py:676)."""
