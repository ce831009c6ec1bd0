"""Build libphc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m puffer_phc_b200.build [--force] [--verbose]

Flags that matter for parity: -fmad=false (no FMA contraction: every fp32 op individually rounded, like
torch eager), default precise division / sqrt, no --use_fast_math.  -lineinfo keeps ncu's source page usable.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libphc_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-O2", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or install the CUDA toolkit)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "phc_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = None, extra_flags=(), only=None) -> str:
    """``out`` / ``extra_flags``: tuning builds into another path (e.g. -DST_WHINT=200), selected at run time with PHC_B200_LIB;
    ``only``: restrict the build to these source files of csrc/ (test builds of a single kernel)."""
    target = out or LIB_PATH
    if out is None and not force and not needs_build():
        return LIB_PATH
    os.makedirs(os.path.dirname(target), exist_ok=True)
    extra = os.environ.get("PHC_NVCC_EXTRA", "").split() + list(extra_flags)      # e.g. -DST_MIN_CTAS=2 for tuning experiments
    srcs = sources() if only is None else [os.path.join(CSRC, f) for f in only]
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", target, *srcs]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libphc_b200.so")
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
