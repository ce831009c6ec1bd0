"""Drop-in for the reference's Cython module ``c_gae`` (reference puffer_phc/c_gae.pyx:11-32; call site
puffer_phc/clean_pufferl/core.py:249): ``compute_gae(dones, values, rewards, gamma, gae_lambda)``.

* numpy float32 arrays in -> new numpy float32 array out (signature parity: H2D, CUDA kernel, D2H);
* CUDA tensors in -> CUDA tensor out, zero-copy (what a device-resident rollout buffer should use).

Flat serial-scan semantics of the reference are kept exactly (see csrc/gae.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _ffi


def compute_gae_cuda(dones: torch.Tensor, values: torch.Tensor, rewards: torch.Tensor, gamma: float, gae_lambda: float,
                     out: torch.Tensor = None, mode: int = 0) -> torch.Tensor:
    lib = _ffi.load()
    _ffi.require_cuda(dones, values, rewards)
    d, v, r = (x.to(torch.float32).contiguous().view(-1) for x in (dones, values, rewards))
    L = r.numel()
    if not (d.numel() == L and v.numel() == L):
        raise ValueError("compute_gae: dones, values and rewards must have the same length")
    adv = torch.empty(L, dtype=torch.float32, device=r.device) if out is None else out
    with _ffi.on_device(r.device):
        _ffi.check(lib.phc_gae(_ffi.ptr(d), _ffi.ptr(v), _ffi.ptr(r), L, float(gamma), float(gae_lambda), _ffi.ptr(adv), int(mode),
                               _ffi.stream_ptr()), "compute_gae")
    return adv


def compute_gae(dones, values, rewards, gamma, gae_lambda):
    if torch.is_tensor(rewards):
        return compute_gae_cuda(dones, values, rewards, gamma, gae_lambda)
    if not torch.cuda.is_available():
        raise RuntimeError("puffer_phc_b200.c_gae: no CUDA device -- the CUDA kernel is the only implementation")
    dev = torch.device("cuda", torch.cuda.current_device())
    d, v, r = (torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev, non_blocking=True) for x in (dones, values, rewards))
    return compute_gae_cuda(d, v, r, gamma, gae_lambda).cpu().numpy()
