"""Parity report: run the CUDA path and the C oracle on the BASELINE configs and print the largest deviations.

    python profiles/tools/parity_report.py > gpurun_out/parity_report.json      (on a GPU box)

For every floating-point output: max abs error, max error relative to max(|ref|, floor) and the number of elements outside
1e-5*|ref| + 2e-6*max(1, row max); for every integer / flag output: the number of mismatches (must be 0).
Test infrastructure: imports the oracle.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co                                   # noqa: E402
from puffer_phc_b200 import synth                                   # noqa: E402
from puffer_phc_b200.c_gae import compute_gae_cuda                  # noqa: E402
from puffer_phc_b200.fused_step import FusedStep, StepConfig        # noqa: E402
from puffer_phc_b200.motion_lib import MotionLibSMPL, STATE_KEYS    # noqa: E402

DEV = "cuda:0"


def npy(t):
    return t.detach().cpu().numpy()


def fstats(got, want):
    got, want = got.astype(np.float64), want.astype(np.float64)
    err = np.abs(got - want)
    floor = 2e-6 * (np.maximum(1.0, np.abs(want).max(axis=-1, keepdims=True)) if want.ndim >= 2 else 1.0)
    bad = err > 1e-5 * np.abs(want) + floor
    return {"max_abs_err": float(err.max()), "max_err_over_tol": float((err / (1e-5 * np.abs(want) + floor)).max()),
            "outside_tolerance": int(bad.sum()), "elements": int(want.size)}


def config(tables_dev, tables_host, N, seed, name):
    lib = MotionLibSMPL.from_tables(tables_dev, device=DEV)
    S = synth.make_env_state(tables_dev, N, seed=seed)
    fs = FusedStep(lib, N, StepConfig())
    out = fs(S["body_state"], S["progress"], S["start_time"], S["start_offset"], S["motion_ids"], S["global_offset"], S["dof_force"], S["dof_vel"])
    torch.cuda.synchronize()
    tab = co.Tables(**{k: tables_host[k] for k in co.TABLE_KEYS})
    want = co.step(tab, npy(S["body_state"]), npy(S["progress"]), npy(S["start_time"]), npy(S["start_offset"]), npy(S["motion_ids"]),
                   npy(S["global_offset"]), 1.0 / 30.0, [100.0, 10.0, 0.1, 0.1], [0.5, 0.3, 0.1, 0.1], np.full(24, 0.25, np.float32),
                   dof_force=npy(S["dof_force"]), dof_vel=npy(S["dof_vel"]))
    rep = {"envs": N, "resets": int(want["reset"].sum()), "terminations": int(want["terminated"].sum())}
    for k in ("reset", "terminated"):
        rep[k + "_mismatches"] = int((npy(out[k]) != want[k]).sum())
    for k in ("obs", "reward", "reward_raw"):
        rep[k] = fstats(npy(out[k]), want[k])
    t0 = (S["progress"].float() * torch.tensor(1.0 / 30.0, device=DEV) + S["start_time"]) + S["start_offset"]
    got = lib.get_motion_state(S["motion_ids"], t0, S["global_offset"], debug=True)
    wantms, (i0, i1, bl) = co.motion_state(tab, npy(S["motion_ids"]), npy(t0), npy(S["global_offset"]), debug=True)
    rep["frame_idx0_mismatches"] = int((npy(got["frame_idx0"]) != i0).sum())
    rep["frame_idx1_mismatches"] = int((npy(got["frame_idx1"]) != i1).sum())
    rep["blend_bit_mismatches"] = int((npy(got["blend"]).view(np.uint32) != bl.view(np.uint32)).sum())
    rep["motion_state"] = {k: fstats(npy(got[k]), wantms[k]) for k in STATE_KEYS}
    return name, rep


def main():
    report = {}
    z = np.load(os.path.join(ROOT, "tests", "golden", "cmu_tables.npz"))
    host = {k: z[k] for k in z.files}
    dev = {k: torch.from_numpy(v).to(DEV) for k, v in host.items()}
    name, rep = config(dev, host, 1024, 1, "config1_cmu_clip_1024_envs")
    report[name] = rep
    T = synth.make_motion_library(11313, seed=0, device=DEV)
    hostT = {k: v.cpu().numpy() for k, v in T.items()}
    for N, seed, nm in ((4096, 1, "config2_amass_4096_envs"), (65536, 3, "config4_amass_65536_envs")):
        name, rep = config(T, hostT, N, seed, nm)
        report[name] = rep
    R = synth.make_rollout(4096, 32, seed=2)
    d, v, r = (R[k].numpy() for k in ("dones", "values", "rewards"))
    adv = npy(compute_gae_cuda(*(torch.from_numpy(x).to(DEV) for x in (d, v, r)), 0.98, 0.2))
    report["config3_gae_4096x32"] = {"bit_mismatches_vs_oracle": int((adv.view(np.uint32) != co.gae(d, v, r, 0.98, 0.2).view(np.uint32)).sum())}
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        import c_gae
        report["config3_gae_4096x32"]["bit_mismatches_vs_reference_c_gae"] = int((adv.view(np.uint32) != c_gae.compute_gae(d, v, r, 0.98, 0.2).view(np.uint32)).sum())
    except ImportError:
        pass
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
