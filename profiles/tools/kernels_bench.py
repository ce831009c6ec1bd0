"""Per-kernel micro-benchmark of the stand-alone C-ABI entry points (CUDA events, rotating inputs, 65536 and 4096 envs):
achieved GB/s on the algorithmic bytes of DESIGN.md section 4 against the measured HBM peak.

    python profiles/tools/kernels_bench.py > gpurun_out/kernels.json
    KB_ONCE=1 ncu --set full -k regex:"reward_kernel|gae_" -c 8 python profiles/tools/kernels_bench.py      (two calls per entry point
                                                                                                          at 65536 envs, nothing timed)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from puffer_phc_b200 import synth                                   # noqa: E402
from puffer_phc_b200.c_gae import compute_gae_cuda                  # noqa: E402
from puffer_phc_b200.envs import common                             # noqa: E402
from puffer_phc_b200.motion_lib import MotionLibSMPL                # noqa: E402
from puffer_phc_b200.policies.running_norm import RunningNorm       # noqa: E402

DEV = "cuda:0"
K = dict(k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1, w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fns, iters=60, warm=10):
    for i in range(warm):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fns[i % len(fns)]()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def graph_time(fns, iters=40):
    """GPU-only time per call: the calls are captured in a CUDA graph (no Python / allocator time between launches)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fns[i % len(fns)]()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    T = synth.make_motion_library(11313, seed=0, device=DEV)
    lib = MotionLibSMPL.from_tables(T, device=DEV)
    pk = peak()
    rows = []
    for N in (65536, 4096):
        sets = [synth.make_env_state(T, N, seed=1 + s) for s in range(4)]
        dt = torch.tensor(1.0 / 30.0, device=DEV)
        times = [(S["progress"].float() * dt + S["start_time"]) + S["start_offset"] for S in sets]
        refs = [lib.get_motion_state(S["motion_ids"], t, S["global_offset"]) for S, t in zip(sets, times)]
        st = [S["body_state"][:, :24] for S in sets]
        views = [(s[..., 0:3], s[..., 3:7], s[..., 7:10], s[..., 10:13]) for s in st]
        obs = [torch.randn(N, 934, device=DEV) for _ in range(4)]
        rn = RunningNorm(934).to(DEV)
        rn.update(obs[0])
        prog = [S["progress"] for S in sets]
        pt = [torch.zeros(N, dtype=torch.bool, device=DEV) for _ in sets]
        rb = torch.ones(N, dtype=torch.bool, device=DEV)
        td = torch.full((24,), 0.25, device=DEV)
        roll = [synth.make_rollout(N, 32, seed=2 + s, device=DEV) for s in range(2)]

        def R(i, k):
            return refs[i][k]
        cases = [
            ("get_motion_state (13 outputs)", 6508, [lambda i=i: lib.get_motion_state(sets[i]["motion_ids"], times[i], sets[i]["global_offset"]) for i in range(4)]),
            ("compute_imitation_observations_v6", 2 * 1248 + 28 + 2304, [lambda i=i: common.compute_imitation_observations_v6(views[i][0][:, 0], views[i][1][:, 0], *views[i], R(i, "rg_pos"), R(i, "rb_rot"), R(i, "body_vel"), R(i, "body_ang_vel"), 1, True) for i in range(4)]),
            ("compute_humanoid_observations_smpl_max", 1248 + 1432, [lambda i=i: common.compute_humanoid_observations_smpl_max(*views[i], None, None, True, True, True, False, False) for i in range(4)]),
            ("compute_imitation_reward", 2 * 1248 + 20, [lambda i=i: common.compute_imitation_reward(views[i][0][:, 0], views[i][1][:, 0], *views[i], R(i, "rg_pos"), R(i, "rb_rot"), R(i, "body_vel"), R(i, "body_ang_vel"), K) for i in range(4)]),
            ("compute_humanoid_im_reset", 2 * 288 + 3 + 2, [lambda i=i: common.compute_humanoid_im_reset(rb, prog[i], None, None, views[i][0], R(i, "rg_pos"), pt[i], True, td, False) for i in range(4)]),
            ("RunningNorm.forward [N,934]", 2 * 3736, [lambda i=i: rn(obs[i]) for i in range(4)]),
            ("RunningNorm.update [N,934]", 3736, [lambda i=i: rn.update(obs[i]) for i in range(4)]),
            ("c_gae.compute_gae [N*32]", 16 * 32, [lambda i=i: compute_gae_cuda(roll[i]["dones"], roll[i]["values"], roll[i]["rewards"], 0.98, 0.2) for i in range(2)]),
        ]
        if os.environ.get("KB_ONCE"):
            for name, _, fns in cases:
                fns[0]()
                fns[1]()
            torch.cuda.synchronize()
            break
        for name, bytes_per_env, fns in cases:
            t = timeit(fns)
            try:
                tg = graph_time(fns)
            except Exception as exc:          # numpy-returning or syncing paths cannot be captured
                tg = float("nan")
            gbs = bytes_per_env * N / tg / 1e9
            rows.append({"kernel": name, "envs": N, "us_eager_api": t * 1e6, "us_gpu": tg * 1e6, "algorithmic_bytes_per_env": bytes_per_env,
                         "GBps": gbs, "frac_of_measured_peak": gbs / pk})
            print(f"{N:6d} {name:42s} api {t*1e6:8.1f} us   gpu {tg*1e6:8.1f} us  {gbs:8.1f} GB/s  {100*gbs/pk:5.1f}%", file=sys.stderr)
    print(json.dumps({"peak_GBps": pk, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
