"""NUMA-aware pinned host buffers for the host-buffer entry of the step (``FusedStep.step_host``).

On a multi-socket GPU box a pinned buffer that lives on the wrong socket is copied through the inter-socket link before it reaches the
GPU's PCIe root, and eight ranks that all allocate from node 0 share one socket's memory controllers.  ``cudaHostAlloc`` takes its
pages from the calling thread's NUMA policy, so the helpers below bind the thread (CPU affinity + ``set_mempolicy(MPOL_BIND)``) to the
node the GPU hangs off for the duration of the allocation and the first touch.  Pure host plumbing: no arithmetic lives here.
"""
from __future__ import annotations

import contextlib
import ctypes
import glob
import os
import re
from typing import Dict, List, Optional

import torch

_SYS_set_mempolicy = 238          # x86_64
_MPOL_DEFAULT, _MPOL_BIND = 0, 2


def _parse_cpulist(s: str) -> List[int]:
    out: List[int] = []
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def numa_nodes() -> Dict[int, List[int]]:
    """NUMA node -> CPU ids, from sysfs ({0: all CPUs} when the kernel exposes no topology)."""
    nodes: Dict[int, List[int]] = {}
    for d in glob.glob("/sys/devices/system/node/node[0-9]*"):
        try:
            with open(os.path.join(d, "cpulist")) as f:
                cpus = _parse_cpulist(f.read())
        except OSError:
            continue
        if cpus:
            nodes[int(re.search(r"node(\d+)$", d).group(1))] = cpus
    return nodes or {0: sorted(os.sched_getaffinity(0))}


def gpu_numa_node(index: int) -> Optional[int]:
    """NUMA node of CUDA device ``index`` (sysfs ``numa_node`` of its PCI function), or None when unknown / single-node."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            n = int(f.read().strip())
        return n if n >= 0 else None
    except Exception:
        return None


@contextlib.contextmanager
def on_numa_node(node: Optional[int]):
    """Run the body with the calling thread bound (CPUs and memory policy) to ``node``; a no-op for ``None`` / unknown nodes."""
    nodes = numa_nodes()
    if node is None or node not in nodes or len(nodes) < 2:
        yield
        return
    old = os.sched_getaffinity(0)
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    bound = False
    try:
        os.sched_setaffinity(0, set(nodes[node]) & old or set(nodes[node]))
        bound = libc.syscall(_SYS_set_mempolicy, _MPOL_BIND, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask))) == 0
        yield
    finally:
        if bound:
            libc.syscall(_SYS_set_mempolicy, _MPOL_DEFAULT, None, ctypes.c_ulong(0))
        os.sched_setaffinity(0, old)


def pinned_empty(shape, dtype, device) -> torch.Tensor:
    """A pinned host tensor whose pages live on the NUMA node of CUDA ``device`` (touched here, so the placement is final)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with on_numa_node(gpu_numa_node(idx)):
        t = torch.empty(shape, dtype=dtype).pin_memory()
        t.view(torch.uint8).reshape(-1)[::4096] = 0          # first touch of every page on this node
    return t


def pinned_like(t: torch.Tensor, device) -> torch.Tensor:
    """NUMA-local pinned host copy of ``t`` (any device)."""
    h = pinned_empty(t.shape, t.dtype, device)
    h.copy_(t)
    return h
