"""The motion library: drop-in for the reference's ``MotionLibBase`` / ``MotionLibSMPL``
(reference puffer_phc/motion_lib.py:180-673).

Same constructor (a ``motion_lib_cfg`` namespace), method names, argument meaning, returned dict keys/shapes and
attribute names the reference's callers read (``_motion_lengths``, ``_num_unique_motions``, ``_curr_motion_ids``,
``_motion_data_keys``, ``_termination_history``).

* query side (``get_motion_state`` & co, motion_lib.py:510-673): csrc/motion_state.cu;
* load side (``load_data`` / ``load_motions``, motion_lib.py:190-429, 744-825): the raw clips stay resident in HBM as
  float64 arrays (motion_file.RawClips) and ``load_motions`` is ONE kernel launch (csrc/build_tables.cu) that does the
  local-rotation / forward-kinematics / filtered-velocity / dof-velocity build for every sampled clip and writes the
  concatenated tables plus the packed frame records of the fused step -- no per-clip host loop, no worker processes.
  Host RNG calls (crop start ``random.randint``, heading ``np.random.random``, ``torch.multinomial``) are made in the
  reference's order so the same seeds pick the same crops.  ``fix_height`` needs the un-vendored ``smpl_sim`` mesh
  parser; like the reference without SMPL model files (motion_lib.py:692-694) the height fix is skipped.
* ``from_tables`` hands in ready-made tables in ``load_motions``' layout (synthetic libraries, tests, benchmarks).

All arithmetic runs in ``libphc_b200.so``; random numbers stay on torch's / Python's generators so sampled ids, crops and
phases are bit-identical to the reference's.
"""
from __future__ import annotations

import ctypes as C
import glob
import os.path as osp
import random
from enum import Enum
from typing import Dict, Optional

import numpy as np
import torch

from . import _ffi
from .motion_file import RawClips


class FixHeightMode(Enum):          # motion_lib.py:66-69
    no_fix = 0
    full_fix = 1
    ankle_fix = 2


class MotionlibMode(Enum):          # motion_lib.py:61-63
    file = 1
    directory = 2


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

    def __init__(self, motion_lib_cfg, pack: bool = True):
        """motion_lib.py:181-191.  ``motion_lib_cfg``: namespace with motion_file, device, fix_height, min_length, max_length,
        im_eval, num_thread (ignored: the build is one kernel), step_dt, is_deterministic."""
        self.m_cfg = motion_lib_cfg
        self._sim_fps = 1 / getattr(self.m_cfg, "step_dt", 1 / 30)
        self._device = torch.device(self.m_cfg.device)
        if self._device.type != "cuda":
            raise RuntimeError("puffer_phc_b200.MotionLib: needs a CUDA device (no CPU implementation)")
        self.mesh_parsers = None
        self._pack = pack
        self.packed = None
        self._lib = _ffi.load()
        self.load_data(self.m_cfg.motion_file, min_length=self.m_cfg.min_length, im_eval=self.m_cfg.im_eval)
        self.setup_constants(fix_height=getattr(self.m_cfg, "fix_height", FixHeightMode.no_fix),
                             num_thread=getattr(self.m_cfg, "num_thread", 1))

    # ---- load side -----------------------------------------------------------------------------------
    def load_data(self, motion_file, min_length=-1, im_eval=False):
        """motion_lib.py:190-227: a pkl (or flat PHCMOT01) file, or a directory of one-clip pkl files; clips shorter than
        ``min_length`` dropped, or (``im_eval`` with min_length == -1) sorted longest first."""
        if isinstance(motion_file, RawClips):
            self.mode, raw = MotionlibMode.file, motion_file
        elif osp.isfile(motion_file):
            self.mode, raw = MotionlibMode.file, RawClips.open(motion_file)
        else:
            self.mode = MotionlibMode.directory
            files = glob.glob(osp.join(motion_file, "*.pkl"))
            assert len(files) > 0
            import joblib
            clips = {}
            for fpath in files:                      # load_motion_with_skeleton reads file[key] lazily (motion_lib.py:768-770)
                key = fpath.split("/")[-1].split(".")[0]
                clips[fpath] = joblib.load(fpath)[key]
            raw = RawClips.from_dict(clips)
        if self.mode == MotionlibMode.file:
            n = len(raw)
            if min_length != -1:
                order = [i for i in range(n) if raw.num_frames[i] >= min_length]
                raw = raw if len(order) == n else raw.subset(order)
            elif im_eval:
                order = sorted(range(n), key=lambda i: int(raw.num_frames[i]), reverse=True)     # stable, like sorted() on items
                raw = raw if order == list(range(n)) else raw.subset(order)
        self._raw = raw
        self._motion_data_keys = raw.keys
        self._num_unique_motions = len(raw)
        raw.to_device(self._device)

    @property
    def _motion_data_list(self):
        """The reference's array of clip dicts (views into the concatenated arrays)."""
        return np.array([self._raw.clip(i) for i in range(len(self._raw))], dtype=object)

    def setup_constants(self, fix_height=FixHeightMode.full_fix, num_thread=1):      # motion_lib.py:229-241
        self.fix_height = fix_height
        self.num_thread = max(num_thread, 1)
        self._curr_motion_ids = None
        self._termination_history = torch.zeros(self._num_unique_motions).to(self._device)
        self._success_rate = torch.zeros(self._num_unique_motions).to(self._device)
        self._sampling_history = torch.zeros(self._num_unique_motions).to(self._device)
        self._sampling_prob = torch.ones(self._num_unique_motions).to(self._device) / self._num_unique_motions
        self._sampling_batch_prob = None

    def load_motions(self, skeleton_trees, gender_betas, limb_weights, random_sample=True, start_idx=0, max_len=-1,
                     sample_idxes=None, dedupe: bool = False):
        """motion_lib.py:257-429: one motion per skeleton (env slot), tables concatenated over the slots.

        ``dedupe`` (extension): slots that load the same clip with the same crop and no random heading share ONE copy of the
        frame rows (``length_starts`` of those slots coincide), so a 65536-env evaluation over 11313 clips holds each clip
        once instead of ~6 times.  Query results are unchanged; only ``_motion_aa`` rows then follow the frame tables' layout
        instead of the reference's uncropped concatenation.
        """
        cfg, raw, dev = self.m_cfg, self._raw, self._device
        n = len(skeleton_trees)
        self.num_joints = len(skeleton_trees[0].node_names)
        if self.num_joints != raw.J:
            raise ValueError(f"skeleton has {self.num_joints} joints, the clips {raw.J}")
        if sample_idxes is None or len(sample_idxes) != n:
            if not cfg.is_deterministic and random_sample:
                sample_idxes = torch.multinomial(self._sampling_prob, num_samples=n, replacement=True).to(dev)
            else:
                sample_idxes = torch.remainder(torch.arange(n) + start_idx, self._num_unique_motions).to(dev)
        sample_idxes = torch.as_tensor(sample_idxes).to(dev)
        self._curr_motion_ids = sample_idxes
        idx = sample_idxes.cpu().numpy().astype(np.int64)
        self.curr_motion_keys = self._motion_data_keys[idx]
        self._sampling_batch_prob = self._sampling_prob[self._curr_motion_ids] / self._sampling_prob[self._curr_motion_ids].sum()

        # ---- per-slot crop and heading: host RNG calls in the reference's order (motion_lib.py:773-799) ----
        seq_len = raw.num_frames[idx]
        cap = cfg.max_length
        crop = np.zeros(n, dtype=np.int64)
        kept = seq_len.copy()
        long = np.zeros(n, dtype=bool) if cap == -1 else (seq_len >= cap)
        kept[long] = cap
        randomise = not (cfg.is_deterministic or cfg.im_eval)
        heading = None
        if randomise or (not cfg.is_deterministic and long.any()):
            heading = np.zeros(n, dtype=np.float64) if randomise else None
            for f in range(n):
                if long[f] and not cfg.is_deterministic:
                    crop[f] = random.randint(0, int(seq_len[f]) - cap)
                if randomise:
                    heading[f] = np.pi * (2 * np.random.random() - 1.0)
        if int(kept.min()) < 2:
            raise ValueError("load_motions: every clip needs at least 2 frames (np.gradient / the dof-velocity loop raise in the reference)")

        # ---- which clips get built -------------------------------------------------------------------
        same_tree = all(t is skeleton_trees[0] for t in skeleton_trees)
        dedupe = dedupe and heading is None
        if dedupe:
            tree_no = np.zeros(n, dtype=np.int64)
            if not same_tree:                       # slots only share rows when they also share the skeleton object
                seen = {}
                tree_no = np.array([seen.setdefault(id(t), len(seen)) for t in skeleton_trees], dtype=np.int64)
            _, first, inverse = np.unique(np.stack([idx, crop, tree_no], 1), axis=0, return_index=True, return_inverse=True)
            inverse = inverse.reshape(-1)
        else:
            first, inverse = np.arange(n), np.arange(n)
        b_nf = kept[first]
        b_out = np.zeros(len(first) + 1, dtype=np.int64)
        np.cumsum(b_nf, out=b_out[1:])
        F = int(b_out[-1])
        tiles = np.zeros(len(first) + 1, dtype=np.int64)
        np.cumsum((b_nf + _ffi.BUILD_TILE - 1) // _ffi.BUILD_TILE, out=tiles[1:])
        J = self.num_joints
        if same_tree:
            lt = skeleton_trees[0].local_translation.to(torch.float32).reshape(1, J, 3)
        else:
            lt = torch.stack([skeleton_trees[int(i)].local_translation.to(torch.float32) for i in first])
        parents = torch.as_tensor(np.asarray(skeleton_trees[0].parent_indices), dtype=torch.int32)

        def up(a, dtype):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype)

        d = raw.to_device(dev)
        meta = {"in_start": up(raw.starts[idx[first]] + crop[first], torch.int64), "num_frames": up(b_nf, torch.int64),
                "out_start": up(b_out[:-1], torch.int64), "fps": up(raw.fps[idx[first]], torch.int32),
                "tile_prefix": up(tiles, torch.int64), "parents": parents.to(dev), "lt": lt.contiguous().to(dev),
                "heading": None if heading is None else up(heading[first], torch.float64)}
        for attr in ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "grvs", "gravs", "packed"):     # motion_lib.py:268-279
            if hasattr(self, attr):
                delattr(self, attr)
        self.packed = None
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)   # noqa: E731
        self.gts, self.grs, self.lrs = new(F, J, 3), new(F, J, 4), new(F, J, 4)
        self.gvs, self.gavs, self.dvs = new(F, J, 3), new(F, J, 3), new(F, J - 1, 3)
        packed = new(F, 312) if (self._pack and J == 24) else None
        bi = _ffi.BuildIn(d["pose_quat_global"].data_ptr(), d["root_trans"].data_ptr(), meta["in_start"].data_ptr(),
                          meta["num_frames"].data_ptr(), meta["out_start"].data_ptr(), meta["fps"].data_ptr(),
                          meta["tile_prefix"].data_ptr(), meta["parents"].data_ptr(), meta["lt"].data_ptr(),
                          0 if same_tree else J * 3, None if heading is None else meta["heading"].data_ptr(),
                          len(first), int(tiles[-1]), J)
        bo = _ffi.BuildOut(*[t.data_ptr() for t in (self.gts, self.grs, self.lrs, self.gvs, self.gavs, self.dvs)],
                           None if packed is None else packed.data_ptr())
        with _ffi.on_device(dev):
            _ffi.check(self._lib.phc_build_motion_tables(C.byref(bi), C.byref(bo), _ffi.stream_ptr()), "phc_build_motion_tables")
        self.packed = packed
        self.grvs, self.gravs = self.gvs[:, 0], self.gavs[:, 0]            # global_root_(angular_)velocity = body 0 (:408-409)

        # ---- _motion_aa: the reference appends each slot's UNCROPPED pose_aa (motion_lib.py:381) ----------
        if dedupe:                          # one segment per built clip, aligned with the frame rows
            seg_slot, seg_src, seg_len, c_lo = first, raw.starts[idx[first]] + crop[first], b_nf, np.zeros(len(first), np.int64)
        else:                               # one segment per slot: the whole clip, heading applied inside the crop window
            seg_slot, seg_src, seg_len, c_lo = np.arange(n), raw.starts[idx], seq_len, crop
        seg_dst = np.zeros(len(seg_len) + 1, dtype=np.int64)
        np.cumsum(seg_len, out=seg_dst[1:])
        aa_w = int(raw.pose_aa.shape[1])
        has_beta = raw.has_beta[idx]
        self._motion_aa = new(int(seg_dst[-1]), aa_w)
        aa = {"src": up(seg_src, torch.int64), "dst": up(seg_dst, torch.int64), "lo": up(c_lo, torch.int64),
              "hi": up(c_lo + kept[seg_slot], torch.int64)}
        with _ffi.on_device(dev):
            _ffi.check(self._lib.phc_build_motion_aa(
                d["pose_aa"].data_ptr(), aa_w, aa["src"].data_ptr(), aa["dst"].data_ptr(), len(seg_len), int(seg_dst[-1]),
                None if heading is None else meta["heading"].data_ptr(), aa["lo"].data_ptr(), aa["hi"].data_ptr(),
                self._motion_aa.data_ptr(), _ffi.stream_ptr()), "phc_build_motion_aa")
        for i in np.nonzero(~has_beta[seg_slot])[0]:                        # motion_lib.py:384-385: zeros for clips without beta
            self._motion_aa[int(seg_dst[i]):int(seg_dst[i + 1])] = 0

        # ---- per-slot metadata: Python double arithmetic, then float32 tensors (motion_lib.py:372-403) ----
        fps = raw.fps[idx].astype(np.float64)
        self._motion_lengths = torch.tensor(1.0 / fps * (kept - 1), device=dev, dtype=torch.float32)
        self._motion_fps = torch.tensor(fps, device=dev, dtype=torch.float32)
        self._motion_dt = torch.tensor(1.0 / fps, device=dev, dtype=torch.float32)
        self._motion_num_frames = torch.tensor(kept, device=dev)
        gb = torch.as_tensor(gender_betas).detach().cpu().to(torch.float32)
        bodies = gb.clone()
        bodies[torch.from_numpy(~has_beta)] = 0                                  # torch.zeros(17) without beta
        self._motion_bodies = bodies.to(dev)
        lw = limb_weights.detach().cpu().numpy() if torch.is_tensor(limb_weights) else np.array(limb_weights)
        self._motion_limb_weights = torch.tensor(lw, device=dev, dtype=torch.float32)
        self._num_motions = n
        self.length_starts = torch.from_numpy(b_out[:-1][inverse]).to(dev)      # :416-419 (shared rows when de-duplicated)
        self.motion_ids = torch.arange(n, dtype=torch.long, device=dev)
        self.num_bodies = J
        self._heading = heading
        self._crop_start = crop
        self._ctables = self._make_ctables()
        return None

    @property
    def one_hot_motions(self):
        """motion_lib.py:310-312 ("Testing for obs_v5"): built on demand -- [slots, clips] int64 is 5.9 GB at 65536 x 11313."""
        return torch.nn.functional.one_hot(self._curr_motion_ids, num_classes=self._num_unique_motions).to(self._device)

    # ------------------------------------------------------------------------------------------------
    @classmethod
    def from_tables(cls, tables: Dict[str, torch.Tensor], device=None, motion_data_keys=None, sim_fps: float = 30.0,
                    pack: bool = True) -> "MotionLibBase":
        """Ready-made tables in ``load_motions``' layout (synthetic libraries, tests, benchmarks)."""
        self = object.__new__(cls)
        dev = torch.device(device) if device is not None else tables["gts"].device
        if dev.type != "cuda":
            raise RuntimeError("puffer_phc_b200.MotionLib: tables must live on a CUDA device (no CPU implementation)")
        self.m_cfg, self._raw, self.mesh_parsers = None, None, None
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
        self._motion_data_keys = motion_data_keys
        self.setup_constants(FixHeightMode.no_fix, 1)
        self._curr_motion_ids = self.motion_ids.clone()
        self._sampling_batch_prob = self._sampling_prob[self._curr_motion_ids] / self._sampling_prob[self._curr_motion_ids].sum()
        self.packed = None
        self._pack = pack
        self._lib = _ffi.load()
        self._ctables = self._make_ctables()
        if pack:
            self.pack()
        return self

    def _make_ctables(self) -> _ffi.MotionTables:
        vals = [getattr(self, _TABLE_ATTR[f]).data_ptr() for f in _ffi.TABLE_FIELDS[:-1]]
        packed = None if self.packed is None else self.packed.data_ptr()
        # the pair tables of the fused step belong to the packed layout: (re)built whenever the frame tables are, for the package's
        # current reference-device flavour; ctables_for(flavour) returns a descriptor whose pair tables match another flavour
        self._pair = {}
        ct = _ffi.MotionTables(*vals, packed, None, None, 0, int(self.gts.shape[0]), self._num_motions)
        if packed is not None and self.grs.shape[1] == 24:
            ct = self._with_pair_tables(ct, _ffi.ref_device())
        return ct

    def _with_pair_tables(self, ct: _ffi.MotionTables, flavour: int) -> _ffi.MotionTables:
        if flavour not in self._pair:
            F = int(self.gts.shape[0])
            aux = torch.empty((F, 24, 2), dtype=torch.float32, device=self._device)
            flags = torch.empty(F, dtype=torch.uint8, device=self._device)
            with _ffi.on_device(self._device):
                _ffi.check(self._lib.phc_build_pair_aux(C.byref(ct), flavour, _ffi.ptr(aux), _ffi.ptr(flags), _ffi.stream_ptr()),
                           "phc_build_pair_aux")
            self._pair[flavour] = (aux, flags)
        aux, flags = self._pair[flavour]
        out = _ffi.MotionTables()
        C.memmove(C.byref(out), C.byref(ct), C.sizeof(ct))
        out.pair_aux, out.pair_flags, out.pair_device = aux.data_ptr(), flags.data_ptr(), flavour
        return out

    def ctables_for(self, flavour: int) -> _ffi.MotionTables:
        """The table descriptor with pair tables built for ``flavour`` (PHC_REF_DEVICE_*); built on first use, then cached."""
        if self.packed is None or self._ctables.pair_device == flavour and self._ctables.pair_aux:
            return self._ctables
        return self._with_pair_tables(self._ctables, flavour)

    def pack(self) -> torch.Tensor:
        """Build the B200 frame layout: one contiguous 1248-byte record (gts|grs|gvs|gavs) per frame."""
        F = int(self.gts.shape[0])
        packed = torch.empty((F, 312), dtype=torch.float32, device=self._device)
        with _ffi.on_device(self._device):
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

    def sample_time_interval(self, motion_ids, truncate_time=None, cpu_division=None):
        """motion_lib.py:526-535.  ``cpu_division=True`` selects the reference's CPU rounding (true division), ``False`` what
        the reference computes when it runs on CUDA (multiplication by float(1.0 / (1/30)) = 30.0f); ``None`` follows the package's
        reference-device setting (``puffer_phc_b200.set_reference_device``, default CUDA)."""
        phase = torch.rand(motion_ids.shape, device=self._device)
        motion_len = self._motion_lengths[motion_ids]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len -= truncate_time
        return self.time_interval_from_phase(phase, motion_len, cpu_division)

    def time_interval_from_phase(self, phase, motion_len, cpu_division=None):
        _ffi.require_cuda(phase, motion_len)
        if cpu_division is None:
            cpu_division = _ffi.ref_device() == _ffi.REF_CPU
        phase, motion_len = phase.contiguous().float(), motion_len.contiguous().float()
        out = torch.empty_like(phase)
        with _ffi.on_device(self._device):
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
        with _ffi.on_device(self._device):
            _ffi.check(self._lib.phc_motion_state(C.byref(self._ctables), _ffi.ptr(ids), _ffi.ptr(times), _ffi.ptr(off), B,
                                                  C.byref(so), _ffi.ref_device(), _ffi.stream_ptr()), "phc_motion_state")
        out.update(dbg)
        return out

    def get_root_pos_smpl(self, motion_ids, motion_times):   # motion_lib.py:628-653
        return self.get_motion_state(motion_ids, motion_times, offset=None, keys=("root_pos",))

    def _calc_frame_blend(self, time, len, num_frames, dt):   # noqa: A002  (reference signature, motion_lib.py:655-665)
        """-> ``(frame_idx0, frame_idx1, blend)`` for per-element ``time``, motion ``len``, ``num_frames`` and frame ``dt``."""
        _ffi.require_cuda(time, len, num_frames, dt)
        n = time.numel()
        t, ln, dtt = (x.to(torch.float32).contiguous().view(-1) for x in (time, len, dt))
        nf = num_frames.to(torch.int64).contiguous().view(-1)
        if not (ln.numel() == n and nf.numel() == n and dtt.numel() == n):
            raise ValueError("_calc_frame_blend: time, len, num_frames and dt must have the same number of elements")
        i0 = torch.empty(n, dtype=torch.int64, device=time.device)
        i1 = torch.empty(n, dtype=torch.int64, device=time.device)
        bl = torch.empty(n, dtype=torch.float32, device=time.device)
        with _ffi.on_device(time.device):
            _ffi.check(self._lib.phc_frame_blend(_ffi.ptr(t), _ffi.ptr(ln), _ffi.ptr(nf), _ffi.ptr(dtt), n, _ffi.ptr(i0), _ffi.ptr(i1),
                                                 _ffi.ptr(bl), _ffi.stream_ptr()), "_calc_frame_blend")
        return i0.view(time.shape), i1.view(time.shape), bl.view(time.shape)

    def _get_num_bodies(self):                                # motion_lib.py:667-668
        return self.num_bodies

    # ---- sampling weights (PMCP hard-negative mining; motion_lib.py:454-508).  Plain bookkeeping on _sampling_prob ------------
    def _key_indexes(self, failed_keys):
        all_keys = self._motion_data_keys.tolist()
        return [all_keys.index(k) for k in failed_keys]

    def _uniform_sampling(self):
        self._sampling_prob = torch.ones(self._num_unique_motions).to(self._device) / self._num_unique_motions

    def update_hard_sampling_weight(self, failed_keys):
        """Train only on the failed sequences (uniform over them); uniform over everything when none failed."""
        if len(failed_keys) > 0:
            idx = self._key_indexes(failed_keys)
            self._sampling_prob[:] = 0
            self._sampling_prob[idx] = 1 / len(idx)
            print(f"Auto PMCP: training on only {len(failed_keys)} seqs")
        else:
            self._uniform_sampling()

    def update_soft_sampling_weight(self, failed_keys):
        """Train mostly on the failed sequences: their termination count goes up and the sampling probability follows it."""
        if len(failed_keys) > 0:
            self._termination_history[self._key_indexes(failed_keys)] += 1
            self.update_sampling_prob(self._termination_history)
            print(f"Auto PMCP: training mostly on {len(self._sampling_prob.nonzero())} seqs")
        else:
            self._uniform_sampling()

    def update_sampling_prob(self, termination_history):
        if len(termination_history) == len(self._termination_history) and termination_history.sum() > 0:
            self._sampling_prob[:] = termination_history / termination_history.sum()
            self._termination_history = termination_history
            return True
        return False


class MotionLibSMPL(MotionLibBase):
    """Name the reference's callers import (motion_lib.py:676)."""
