"""CPU check of the PRODUCT's kernel math: puffer_phc_b200/csrc/phc_math.cuh + phc_body.cuh compiled for the
host (tests/host_math_harness.cpp) and replayed per env, against the golden vectors from the reference.
(The CUDA kernels themselves are tested by the -m gpu tests; this catches math/order bugs without a GPU.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_close, assert_equal
from puffer_phc_b200 import _ffi

K = [100.0, 10.0, 0.1, 0.1]
W = [0.5, 0.3, 0.1, 0.1]
EVAL_MASK = sum(1 << j for j in range(24) if j not in (4, 8, 18, 23))


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "host_math_harness.cpp")
    out = os.path.join(ROOT, "tests", "_build", "libhost_harness.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                           "-I/usr/local/cuda/include", "-o", out, src])
    return C.CDLL(out)


def _p(a):
    return C.c_void_p(None if a is None else a.ctypes.data)


def _tables(T):
    keep = {k: np.ascontiguousarray(T[k]) for k in T}
    mt = _ffi.MotionTables(*[keep[k].ctypes.data if k in keep else None for k in _ffi.TABLE_FIELDS[:-1]], None,
                           keep["gts"].shape[0], keep["motion_len"].shape[0])
    return mt, keep


def _run(harness, T, S, term, mask, use_mean, power):
    mt, keep = _tables(T)
    N = S["in_progress"].shape[0]
    bs = np.ascontiguousarray(S["in_body_state"]).reshape(N, -1)
    arrs = dict(bs=bs, prog=np.ascontiguousarray(S["in_progress"]), st=S["in_start_time"], so=S["in_start_offset"],
                ids=S["in_motion_ids"], go=np.ascontiguousarray(S["in_global_offset"]),
                df=np.ascontiguousarray(S["in_dof_force"]), dv=np.ascontiguousarray(S["in_dof_vel"]),
                td=np.full(24, term, np.float32))
    sin = _ffi.StepIn(_p(arrs["bs"]), bs.shape[1], _p(arrs["prog"]), _p(arrs["st"]), _p(arrs["so"]), _p(arrs["ids"]), _p(arrs["go"]),
                      _p(arrs["df"]) if power else None, _p(arrs["dv"]) if power else None, _p(arrs["td"]), None, None, N)
    cfg = _ffi.StepCfg(float(np.float32(1.0 / 30.0)), (C.c_float * 4)(*K), (C.c_float * 4)(*W), float(np.float32(0.0005)),
                       mask, 1, int(use_mean), 1e-5, 10.0, 0)
    rw = 5 if power else 4
    obs, rew, raw = np.zeros((N, 934), np.float32), np.zeros(N, np.float32), np.zeros((N, rw), np.float32)
    rs, tm = np.zeros(N, np.uint8), np.zeros(N, np.uint8)
    sout = _ffi.StepOut(_p(obs), 934, None, _p(rew), _p(raw), rw, _p(rs), _p(tm), None, 0, None, None)
    assert harness.harness_step(C.byref(mt), C.byref(sin), C.byref(cfg), C.byref(sout)) == 0
    return obs, rew, raw, rs.astype(bool), tm.astype(bool)


@pytest.mark.parametrize("tn,sn", [("cmu_tables", "cmu_step"), ("synth_tables", "synth_step")])
def test_kernel_math_vs_reference(harness, golden, tn, sn):
    T, S = golden[tn], golden[sn]
    obs, rew, raw, rs, tm = _run(harness, T, S, 0.25, 0xFFFFFF, False, True)
    assert_close(obs, S["obs"], what="obs", row_scale=True)
    assert_close(rew, S["reward"], what="reward")
    assert_close(raw, S["reward_raw"], what="reward_raw")
    assert_equal(rs, S["reset_train"], "reset")
    assert_equal(tm, S["terminated_train"], "terminated")
    obs, rew, raw, rs, tm = _run(harness, T, S, 0.5, EVAL_MASK, True, False)
    assert_equal(rs, S["reset_eval"], "reset eval")
    assert_equal(tm, S["terminated_eval"], "terminated eval")
    assert_close(rew, S["reward_nopower"], what="reward, no power term")


@pytest.mark.parametrize("tn,sn", [("cmu_tables", "cmu_step"), ("synth_tables", "synth_step")])
def test_slerp_expmap_vs_reference(harness, golden, tn, sn):
    T, S = golden[tn], golden[sn]
    mt, keep = _tables(T)
    ids, times = S["in_motion_ids"], S["t1"]
    B = ids.shape[0]
    dof, rot = np.zeros((B, 69), np.float32), np.zeros((B, 24, 4), np.float32)
    i0, i1, bl = np.zeros(B, np.int64), np.zeros(B, np.int64), np.zeros(B, np.float32)
    harness.harness_dof_pos(C.byref(mt), _p(ids), _p(times), C.c_int64(B), _p(dof), _p(rot), _p(i0), _p(i1), _p(bl))
    assert_equal(i0, S["t1_idx0"], "idx0")
    assert_equal(i1, S["t1_idx1"], "idx1")
    assert_equal(bl.view(np.uint32), S["t1_blend"].view(np.uint32), "blend bits")
    assert_close(dof, S["t1_dof_pos"], what="dof_pos")
    assert_close(rot, S["t1_rb_rot"], what="rb_rot")
