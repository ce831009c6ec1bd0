"""Motion table build (row f4) benchmark: ``MotionLibSMPL.load_motions`` over an AMASS-shaped synthetic raw library that is
resident in HBM (11313 clips, ~2.5 M frames, float64 like the pkl format), CUDA-event timed, against the algorithmic bytes of
DESIGN.md section 4 and the measured HBM peak; the C oracle (single core) on a bounded sample is the CPU baseline.

    python profiles/tools/loader_bench.py > gpurun_out/loader.json
"""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from puffer_phc_b200 import synth                                   # noqa: E402
from puffer_phc_b200.motion_file import RawClips                    # noqa: E402
from puffer_phc_b200.motion_lib import MotionLibSMPL                # noqa: E402
from puffer_phc_b200.skeleton import SkeletonTree                   # noqa: E402

DEV = "cuda:0"
BYTES_IN = 24 * 4 * 8 + 3 * 8 + 72 * 8            # pose_quat_global + root_trans + pose_aa, float64
BYTES_OUT = (72 + 96 + 96 + 72 + 72 + 69 + 72) * 4  # gts grs lrs gvs gavs dvs motion_aa
BYTES_PACKED = 312 * 4


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    T = synth.make_motion_library(11313, seed=0, device=DEV)
    nf = T["num_frames"].cpu().numpy()
    fps = np.round(1.0 / T["motion_dt"].cpu().numpy()).astype(np.int32)
    raw = RawClips.from_device([f"c{i}" for i in range(len(nf))], nf, fps, T["gts"][:, 0].double().contiguous(),
                               T["motion_aa"].double().contiguous(), T["grs"].double().contiguous())
    parents = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22]
    rng = np.random.default_rng(0)
    sk = SkeletonTree([f"b{j}" for j in range(24)], np.array(parents, np.int32), rng.normal(0, 0.15, (24, 3)).astype(np.float32))
    del T
    torch.cuda.empty_cache()
    cfg = SimpleNamespace(motion_file=raw, device=DEV, min_length=-1, max_length=300, im_eval=False, is_deterministic=True, step_dt=1 / 30)
    pk = peak()
    rows = []
    for slots, dedupe, pack in ((11313, False, True), (4096, False, True), (65536, True, True)):
        lib = MotionLibSMPL(cfg, pack=pack)
        idx = torch.arange(slots) % 11313
        kw = dict(skeleton_trees=[sk] * slots, gender_betas=torch.zeros(slots, 17), limb_weights=np.zeros((slots, 10)),
                  sample_idxes=idx, dedupe=dedupe)
        lib.load_motions(**kw)
        torch.cuda.synchronize()
        ev, wall = [], []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            lib.load_motions(**kw)
            b.record()
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
            ev.append(a.elapsed_time(b) * 1e-3)
        frames = int(lib.gts.shape[0])
        aa_rows = int(lib._motion_aa.shape[0])
        # kernel-only: the two launches, no host bookkeeping in between (events around back-to-back re-launches via the ABI are
        # what load_motions does last; the device time of the call is dominated by them, the wall time by the Python bookkeeping)
        alg = frames * (BYTES_IN - 72 * 8 + BYTES_OUT - 72 * 4 + (BYTES_PACKED if pack else 0)) + aa_rows * (72 * 8 + 72 * 4)
        rows.append({"slots": slots, "dedupe": dedupe, "frames_built": frames, "device_ms_best": min(ev) * 1e3, "wall_ms_best": min(wall) * 1e3,
                     "frames_per_s": frames / min(ev), "algorithmic_bytes": alg, "gbs": alg / min(ev) / 1e9, "frac_of_measured_peak": alg / min(ev) / 1e9 / pk})
        del lib
        torch.cuda.empty_cache()

    # CPU baseline: the C oracle (one core) on a bounded sample of clips
    from oracle import c_oracle as co
    d = raw.to_device(DEV)
    sample = list(range(0, 11313, 177))
    t0 = time.perf_counter()
    fr = 0
    for i in sample:
        a, b = int(raw.starts[i]), int(raw.starts[i + 1])
        co.build_clip(d["pose_quat_global"][a:b].cpu().numpy(), d["root_trans"][a:b].cpu().numpy(), np.array(parents), sk.local_translation.numpy(), int(raw.fps[i]))
        fr += b - a
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"peak_gbs": pk, "rows": rows, "cpu_oracle": {"clips": len(sample), "frames": fr, "seconds": cpu_s, "frames_per_s": fr / cpu_s, "cores": 1}}, indent=1))


if __name__ == "__main__":
    main()
