"""Role-skew stress test of the fused step's mbarrier pipeline (advisor finding of round 1: a latent race on the tile barrier that
only the relative speed of the warp roles kept from triggering).

The kernel is rebuilt (on the GPU box, nvcc is part of the image) with -DST_STRESS_DELAY=1/2/3, which makes role B, role A or the
writer warps sleep a pseudo-random few microseconds in every iteration; outputs, normalised outputs, moments and metrics must stay
BIT-IDENTICAL to the production build under any skew."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Proxy:
    """The production library with phc_step_fused swapped for the stress build's."""

    def __init__(self, base, stress):
        self._base, self._stress = base, stress

    def __getattr__(self, name):
        return getattr(self._stress if name == "phc_step_fused" else self._base, name)


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.isfile("/usr/local/cuda/bin/nvcc"), reason="needs nvcc on the GPU box")
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_outputs_are_bit_identical_under_role_skew(mode):
    from puffer_phc_b200 import _ffi, build, synth
    from puffer_phc_b200.fused_step import FusedStep, StepConfig
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    from puffer_phc_b200.policies.running_norm import RunningNorm
    out = os.path.join(ROOT, "tests", "_build", f"libphc_stress{mode}.so")
    build.build(out=out, extra_flags=[f"-DST_STRESS_DELAY={mode}"], only=("step_fused.cu", "errors.cu"))
    stress = C.CDLL(out)
    stress.phc_step_fused.argtypes = _ffi.load().phc_step_fused.argtypes
    stress.phc_step_fused.restype = C.c_int
    T = synth.make_motion_library(200, seed=3, device=DEV, other_fps_fraction=0.3, freeze_every=7)
    lib = MotionLibSMPL.from_tables(T, device=DEV)
    N = 20003                                   # ~17 iterations per CTA, ragged tail
    S = synth.make_env_state(T, N, seed=31)
    args = [S[k] for k in ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")]
    res = []
    for use_stress in (False, True):
        rms = RunningNorm(934).to(DEV)
        rms.running_mean.fill_(0.1); rms.running_var.fill_(1.3)
        fs = FusedStep(lib, N, StepConfig(ref_device="cpu"), rms=rms, normalize=True, accumulate_moments=True, defer_moments=True, metrics=True)
        if use_stress:
            fs.lib = _Proxy(fs.lib, stress)
        for _ in range(2):
            o = fs(*args)
        fs.flush_moments()
        torch.cuda.synchronize()
        res.append({**{k: v.clone() for k, v in o.items()}, "stats": fs.stats.clone()})
    a, b = res
    for k in a:
        x, y = a[k], b[k]
        if x.dtype == torch.float32:
            x, y = x.view(torch.int32), y.view(torch.int32)
        elif x.dtype == torch.float64:
            x, y = x.view(torch.int64), y.view(torch.int64)
        assert torch.equal(x, y), f"{k} differs under -DST_STRESS_DELAY={mode}"
    assert float(a["stats"][0]) == 2 * N and float(a["stats"][1 + 2 * 934]) == 2 * N
