"""Raw motion clips: the reference's on-disk format and a flat binary twin that loads without unpickling.

The reference stores a library as a joblib-compressed pickle ``{key: {"root_trans_offset": [T,3] f64 tensor,
"pose_aa": [T,72] f64, "pose_quat_global": [T,24,4] f64, "beta": [16], "gender": str, "fps": int}}``
(reference scripts/convert_amass_data.py:186-205) and keeps it as a Python list of dicts on the host
(motion_lib.py:190-227); every ``load_motions`` then walks the sampled clips one by one on the CPU.

Here the clips are concatenated once into three float64 arrays that stay RESIDENT IN HBM (an AMASS-sized
library is ~5 GB of 180 GB): ``load_motions`` turns into one kernel launch that reads the sampled clips where
they lie (csrc/build_tables.cu).  ``RawClips.save`` / ``RawClips.load`` write / memory-map the same arrays as
a flat little-endian file:

    magic "PHCMOT01" | u64 n_clips | u64 n_frames | u32 J | u32 aa_width | u64 meta_bytes | meta JSON (keys, fps,
    gender, has_beta) padded to 64 B | i64 num_frames[n_clips] | f64 beta[n_clips,16] | pad to 64 B |
    f64 root_trans_offset[F,3] | f64 pose_aa[F,aa_width] | f64 pose_quat_global[F,J,4]
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Optional, Sequence

import numpy as np
import torch

MAGIC = b"PHCMOT01"


def _np(x, dtype=np.float64):
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def _pad64(n: int) -> int:
    return (-n) % 64


class RawClips:
    """Concatenated raw clips (host numpy, optionally mirrored on a CUDA device)."""

    def __init__(self, keys: Sequence[str], num_frames, fps, root_trans, pose_aa, pose_quat_global, beta=None, gender=None,
                 has_beta=None):
        self.keys = np.array(list(keys))
        self.num_frames = np.ascontiguousarray(num_frames, dtype=np.int64)
        self.fps = np.ascontiguousarray(fps, dtype=np.int32)
        self.starts = np.zeros(len(self.num_frames) + 1, dtype=np.int64)
        np.cumsum(self.num_frames, out=self.starts[1:])
        self.root_trans, self.pose_aa, self.pose_quat_global = root_trans, pose_aa, pose_quat_global
        n = len(self.num_frames)
        self.beta = np.zeros((n, 16)) if beta is None else np.ascontiguousarray(beta, dtype=np.float64)
        self.gender = ["neutral"] * n if gender is None else list(gender)
        self.has_beta = np.ones(n, dtype=bool) if has_beta is None else np.ascontiguousarray(has_beta, dtype=bool)
        F = int(self.starts[-1])
        if root_trans.shape[0] != F or pose_aa.shape[0] != F or pose_quat_global.shape[0] != F:
            raise ValueError("RawClips: arrays do not hold sum(num_frames) rows")
        self.device_arrays: Optional[Dict[str, torch.Tensor]] = None

    # ------------------------------------------------------------------------------------------------
    def __len__(self):
        return len(self.num_frames)

    @property
    def J(self) -> int:
        return int(self.pose_quat_global.shape[1])

    @classmethod
    def from_dict(cls, clips: Dict[str, dict]) -> "RawClips":
        """From the unpickled reference format (dict key -> clip dict)."""
        keys = list(clips.keys())
        vals = [clips[k] for k in keys]
        if not vals:
            raise ValueError("RawClips: empty library")
        nf = [int(np.shape(v["pose_quat_global"])[0]) for v in vals]
        return cls(
            keys, nf, [int(v.get("fps", 30)) for v in vals],
            np.concatenate([_np(v["root_trans_offset"]).reshape(n, 3) for v, n in zip(vals, nf)]),
            np.concatenate([_np(v["pose_aa"]).reshape(n, -1) for v, n in zip(vals, nf)]),
            np.concatenate([_np(v["pose_quat_global"]) for v in vals]),
            beta=np.stack([np.resize(_np(v["beta"]).reshape(-1), 16) if "beta" in v else np.zeros(16) for v in vals]),
            gender=[str(v.get("gender", "neutral")) for v in vals],
            has_beta=["beta" in v for v in vals])

    @classmethod
    def from_pkl(cls, path: str) -> "RawClips":
        import joblib                                   # the reference's container (motion_lib.py:193)
        return cls.from_dict(joblib.load(path))

    @classmethod
    def from_device(cls, keys, num_frames, fps, root_trans, pose_aa, pose_quat_global, **kw) -> "RawClips":
        """Clips that already live on a CUDA device as float64 tensors (synthetic libraries, benchmarks): no host copy is kept."""
        arrs = {"root_trans": root_trans, "pose_aa": pose_aa, "pose_quat_global": pose_quat_global}
        for k, t in arrs.items():
            if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise ValueError(f"RawClips.from_device: {k} must be a contiguous float64 CUDA tensor")
        self = cls(keys, num_frames, fps, root_trans, pose_aa, pose_quat_global, **kw)
        self.device_arrays = arrs
        return self

    def subset(self, order: Sequence[int]) -> "RawClips":
        """Clips ``order`` (filtering / sorting of load_data, motion_lib.py:203-218) as a new concatenation."""
        order = np.asarray(order, dtype=np.int64)
        rows = np.concatenate([np.arange(self.starts[i], self.starts[i + 1]) for i in order]) if len(order) else np.zeros(0, np.int64)
        return RawClips(self.keys[order], self.num_frames[order], self.fps[order], self.root_trans[rows], self.pose_aa[rows],
                        self.pose_quat_global[rows], self.beta[order], [self.gender[i] for i in order], self.has_beta[order])

    def clip(self, i: int) -> dict:
        """Clip i in the reference's dict form (views, no copy)."""
        a, b = int(self.starts[i]), int(self.starts[i + 1])
        return {"root_trans_offset": self.root_trans[a:b], "pose_aa": self.pose_aa[a:b], "pose_quat_global": self.pose_quat_global[a:b],
                "beta": self.beta[i], "gender": self.gender[i], "fps": int(self.fps[i])}

    # ------------------------------------------------------------------------------------------------
    def save(self, path: str) -> None:
        meta = json.dumps({"keys": [str(k) for k in self.keys], "fps": [int(x) for x in self.fps], "gender": self.gender,
                           "has_beta": [bool(x) for x in self.has_beta]}).encode("utf-8")
        meta += b" " * _pad64(len(MAGIC) + 8 + 8 + 4 + 4 + 8 + len(meta))
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(struct.pack("<QQIIQ", len(self), int(self.starts[-1]), self.J, int(self.pose_aa.shape[1]), len(meta)))
            f.write(meta)
            f.write(self.num_frames.astype("<i8").tobytes())
            f.write(np.ascontiguousarray(self.beta, dtype="<f8").tobytes())
            f.write(b"\0" * _pad64(f.tell()))
            for arr in (self.root_trans, self.pose_aa, self.pose_quat_global):
                f.write(np.ascontiguousarray(arr, dtype="<f8").tobytes())

    @classmethod
    def load(cls, path: str) -> "RawClips":
        """Memory-map a flat file written by ``save`` (the big arrays are not read until they are uploaded)."""
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{path}: not a PHCMOT01 motion file")
            n, F, J, aa_w, meta_bytes = struct.unpack("<QQIIQ", f.read(32))
            meta = json.loads(f.read(meta_bytes).decode("utf-8"))
            off = f.tell()
        nf = np.fromfile(path, dtype="<i8", count=n, offset=off)
        off += 8 * n
        beta = np.fromfile(path, dtype="<f8", count=n * 16, offset=off).reshape(n, 16)
        off += 8 * 16 * n
        off += _pad64(off)
        arrs = []
        for shape in ((F, 3), (F, aa_w), (F, J, 4)):
            cnt = int(np.prod(shape))
            arrs.append(np.memmap(path, dtype="<f8", mode="r", offset=off, shape=shape) if cnt else np.zeros(shape))
            off += 8 * cnt
        if off != os.path.getsize(path):
            raise ValueError(f"{path}: truncated or trailing bytes ({os.path.getsize(path)} != {off})")
        return cls(meta["keys"], nf, meta["fps"], arrs[0], arrs[1], arrs[2], beta, meta["gender"], meta["has_beta"])

    @classmethod
    def open(cls, path: str) -> "RawClips":
        """Either container, by content."""
        with open(path, "rb") as f:
            head = f.read(8)
        return cls.load(path) if head == MAGIC else cls.from_pkl(path)

    # ------------------------------------------------------------------------------------------------
    def to_device(self, device) -> Dict[str, torch.Tensor]:
        """Upload (once) and keep the three float64 arrays resident on ``device``."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("puffer_phc_b200: raw clips are built into tables on a CUDA device only (no CPU implementation)")
        have = None if self.device_arrays is None else self.device_arrays["root_trans"].device
        if have is None or have.type != "cuda" or (dev.index is not None and have.index != dev.index):
            self.device_arrays = {k: torch.from_numpy(np.ascontiguousarray(getattr(self, k), dtype=np.float64)).to(dev)
                                  for k in ("root_trans", "pose_aa", "pose_quat_global")}
        return self.device_arrays


def convert_pkl(pkl_path: str, out_path: str) -> RawClips:
    """``amass_train_*.pkl`` (joblib) -> flat PHCMOT01 file."""
    raw = RawClips.from_pkl(pkl_path)
    raw.save(out_path)
    return raw
