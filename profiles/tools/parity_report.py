"""Parity report: run the CUDA path and the C oracle on the BASELINE configs and print the largest deviations.

    python profiles/tools/parity_report.py gpurun_out/parity_report.json      (on a GPU box)

Part 1 (vs the C oracle, torch-CPU flavour): for every floating-point output the max abs error, the worst error over tolerance
under the PLAIN rule 1e-5*|ref| + 2e-6 and (observations) under the per-vector floor 2e-6*max(1, |v|) of the tests; for every
integer / flag output the number of mismatches (must be 0).
Part 2 (vs the REFERENCE'S OWN FUNCTIONS, oracle/_ref, executed here on torch-CPU and torch-CUDA with the kernels' ref_device set
to match): mismatch counts of idx0 / idx1 / blend / reset / terminated, the closest margin of any env to the termination and
pass-time thresholds, and the worst float margins under both rules (oracle/device_parity.py).
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
    from oracle import device_parity as dp
    got, want = got.astype(np.float64), want.astype(np.float64)
    err = np.abs(got - want)
    plain = err / (1e-5 * np.abs(want) + 2e-6)
    rep = {"max_abs_err": float(err.max()), "plain_tol_worst": float(plain.max()), "plain_tol_n_over": int((plain > 1).sum()),
           "elements": int(want.size)}
    if want.ndim == 2 and want.shape[1] == 934:
        vf = err / (1e-5 * np.abs(want) + dp.obs_floor(want))
        rep["vector_floor_worst"], rep["vector_floor_n_over"] = float(vf.max()), int((vf > 1).sum())
    return rep


def config(tables_dev, tables_host, N, seed, name):
    import puffer_phc_b200
    puffer_phc_b200.set_reference_device("cpu")         # the C oracle restates torch-CPU rounding
    lib = MotionLibSMPL.from_tables(tables_dev, device=DEV)
    S = synth.make_env_state(tables_dev, N, seed=seed)
    fs = FusedStep(lib, N, StepConfig(ref_device="cpu"))
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
    import contextlib
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    with contextlib.redirect_stdout(sys.stderr):        # the reference's loader prints progress to stdout
        report = build_report()
    text = json.dumps(report, indent=1)
    if out_path:
        with open(out_path, "w") as f:
            f.write(text)
    else:
        print(text)


def build_report():
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
    # ---- part 2: the reference's own functions on this box, both devices ------------------------------------------------
    from oracle import device_parity as dp, ref_runner as rr
    if rr.available():
        ref = {}
        _, Tc = rr.load_cmu("cpu")
        Tcd = {k: v.to(DEV) for k, v in Tc.items()}
        Sc = synth.make_env_state(Tc, 1024, seed=1)
        lib_big = MotionLibSMPL.from_tables(T, device=DEV)
        for flavour in ("cpu", "cuda"):
            ref[f"config1_cmu_clip_1024_envs/{flavour}"] = dp.compare_step(Tcd, Sc, flavour)[0]
            for N, seed, nm in ((4096, 1, "config2_amass_4096_envs"), (65536, 3, "config4_amass_65536_envs")):
                S = {k: v.cpu() for k, v in synth.make_env_state(T, N, seed=seed).items()}
                ref[f"{nm}/{flavour}"] = dp.compare_step(T, S, flavour, ours_lib=lib_big)[0]
            S = {k: v.cpu() for k, v in synth.make_env_state(T, 65536, seed=5).items()}
            ref[f"eval_variant_65536_envs_threshold_0.3/{flavour}"] = dp.compare_step(T, S, flavour, eval_mode=True, ours_lib=lib_big,
                                                                                    eval_distance=0.3)[0]
        with open(os.path.join(rr.REF, "MANIFEST.json")) as f:
            report["vs_reference_own_functions"] = {"reference_files_sha256": json.load(f)["files"], "torch": torch.__version__, "runs": ref}
    return report


if __name__ == "__main__":
    main()
