"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/phc_b200.h declares,
argument validation works without a GPU, the ctypes structs match the C layout, and the product refuses CPU tensors."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT
from puffer_phc_b200 import _ffi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "phc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.load()
    declared = _declared_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/phc_b200.h but not exported"
    assert set(declared) == set(_ffi.EXPORTS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _ffi.library_path()], text=True)
    exported = set(re.findall(r"\bT (phc_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert lib.phc_version() == 121


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", _ffi.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layout_matches_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "phc_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(phc_view), sizeof(phc_motion_tables), sizeof(phc_motion_state_out),
             sizeof(phc_step_in), sizeof(phc_step_cfg), sizeof(phc_step_out));
      printf("%zu %zu %zu %zu %zu\n", offsetof(phc_step_cfg, power_coef), offsetof(phc_step_cfg, reset_body_mask),
             offsetof(phc_step_cfg, rms_clip), offsetof(phc_step_in, N), offsetof(phc_step_out, moment_partials));
      return 0; }'''
    d = os.path.join(ROOT, "tests", "_build")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "layout.c"), "w").write(src)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "layout"), os.path.join(d, "layout.c")])
    a, b = subprocess.check_output([os.path.join(d, "layout")], text=True).strip().split("\n")
    sizes = [int(x) for x in a.split()]
    assert sizes == [C.sizeof(x) for x in (_ffi.View, _ffi.MotionTables, _ffi.MotionStateOut, _ffi.StepIn, _ffi.StepCfg, _ffi.StepOut)]
    offs = [int(x) for x in b.split()]
    assert offs == [_ffi.StepCfg.power_coef.offset, _ffi.StepCfg.reset_body_mask.offset, _ffi.StepCfg.rms_clip.offset,
                    _ffi.StepIn.N.offset, _ffi.StepOut.moment_partials.offset]


def test_argument_validation_without_gpu():
    lib = _ffi.load()
    assert lib.phc_gae(None, None, None, -1, 0.9, 0.9, None, 0, None) == _ffi.PHC_EINVAL
    assert b"L < 0" in lib.phc_last_error()
    assert lib.phc_gae(None, None, None, 8, 0.9, 0.9, None, 0, None) == _ffi.PHC_EINVAL
    assert lib.phc_gae(None, None, None, 0, 0.9, 0.9, None, 0, None) == _ffi.PHC_OK          # empty input is a no-op
    assert lib.phc_gae(None, None, None, 8, 0.9, 0.9, None, 7, None) == _ffi.PHC_EINVAL
    v = _ffi.View(1, 0, 0)
    assert lib.phc_imitation_obs_v6(*([v] * 10), 4, 24, 0, 1, C.c_void_p(1), 576, None) == _ffi.PHC_EINVAL          # time_steps < 1
    assert lib.phc_imitation_obs_v6(*([v] * 10), 4, 24, 2, 1, C.c_void_p(1), 576, None) == _ffi.PHC_ESHAPE          # row too narrow for 2 steps
    assert lib.phc_imitation_obs_v6(*([v] * 10), 4, 40, 1, 1, C.c_void_p(1), 960, None) == _ffi.PHC_ESHAPE
    assert lib.phc_imitation_obs_v6(*([v] * 10), 0, 24, 1, 1, None, 576, None) == _ffi.PHC_OK
    assert lib.phc_rms_forward(None, 934, None, None, 1e-5, 10.0, 4, 934, None, 934, None) == _ffi.PHC_EINVAL
    assert lib.phc_rms_forward(C.c_void_p(16), 10, C.c_void_p(16), C.c_void_p(16), 1e-5, 10.0, 4, 934, C.c_void_p(16), 934, None) == _ffi.PHC_ESHAPE
    assert lib.phc_step_fused(None, None, None, None, None) == _ffi.PHC_EINVAL
    assert lib.phc_motion_state(None, None, None, None, 3, None, 0, None) == _ffi.PHC_EINVAL
    with pytest.raises(ValueError):
        _ffi.check(_ffi.PHC_EINVAL, "x")
    with pytest.raises(NotImplementedError):
        _ffi.check(_ffi.PHC_EUNSUPPORTED, "x")


def test_product_refuses_cpu_tensors():
    """No CPU fallback: host tensors raise instead of being computed somewhere else."""
    from puffer_phc_b200.envs import common
    from puffer_phc_b200.motion_lib import MotionLibSMPL
    from puffer_phc_b200.policies.running_norm import RunningNorm
    x = torch.zeros(4, 24, 3)
    q = torch.zeros(4, 24, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        common.compute_imitation_reward(x[:, 0], q[:, 0], x, q, x, x, x, q, x, x, dict(k_pos=1, k_rot=1, k_vel=1, k_ang_vel=1, w_pos=1, w_rot=1, w_vel=1, w_ang_vel=1))
    with pytest.raises(RuntimeError, match="CUDA"):
        RunningNorm(934)(torch.zeros(2, 934))
    with pytest.raises(RuntimeError, match="CUDA"):
        MotionLibSMPL.from_tables({"gts": torch.zeros(1, 24, 3)}, device="cpu")
    if not torch.cuda.is_available():
        from puffer_phc_b200 import c_gae
        with pytest.raises(RuntimeError, match="CUDA"):
            c_gae.compute_gae(np.zeros(4, np.float32), np.zeros(4, np.float32), np.zeros(4, np.float32), 0.9, 0.9)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "puffer_phc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle" not in text.replace("# oracle-free", ""), f"{f} mentions the oracle"


def test_shard_ranges_cover_all_envs():
    from puffer_phc_b200.dist import shard_range
    for total, world in ((524288, 8), (65536, 3), (10, 4), (3, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)


def test_synth_library_shapes_and_determinism():
    from puffer_phc_b200 import synth
    a = synth.make_motion_library(20, seed=3, other_fps_fraction=0.5, freeze_every=4)
    b = synth.make_motion_library(20, seed=3, other_fps_fraction=0.5, freeze_every=4)
    F = int(a["num_frames"].sum())
    assert a["gts"].shape == (F, 24, 3) and a["grs"].shape == (F, 24, 4) and a["dvs"].shape == (F, 23, 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert int(a["num_frames"].min()) >= 10 and int(a["num_frames"].max()) <= 300
    assert torch.equal(a["length_starts"], torch.cumsum(a["num_frames"], 0) - a["num_frames"])
    assert torch.allclose(a["grs"].norm(dim=-1), torch.ones(F, 24), atol=1e-5)
    assert bool((a["lrs"][:, [4, 8, 18, 23]] == torch.tensor([0.0, 0.0, 0.0, 1.0])).all())
    S = synth.make_env_state(a, 64, seed=1, bodies_per_env=25)
    assert S["body_state"].shape == (64, 25, 13) and S["progress"].dtype == torch.int16 and S["motion_ids"].dtype == torch.int64
