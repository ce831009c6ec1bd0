"""Is the board's power cap a property of the fused step's arithmetic or of its memory traffic?  Runs (a) the traffic-only kernel of
gather_peak.cu and (b) phc_step_fused back to back for ~3 s each while nvidia-smi samples power / SM clock / throttle reasons.

    python profiles/tools/power_probe.py
"""
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from puffer_phc_b200 import synth                                   # noqa: E402
from puffer_phc_b200.fused_step import FusedStep, StepConfig        # noqa: E402
from puffer_phc_b200.motion_lib import MotionLibSMPL                # noqa: E402
from puffer_phc_b200.policies.running_norm import RunningNorm       # noqa: E402


class Sampler:
    def __init__(self):
        self.lines, self.proc = [], None

    def start(self):
        self.lines = []
        self.proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap",
                                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=lambda: [self.lines.append(l.strip()) for l in self.proc.stdout], daemon=True).start()

    def stop(self):
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(",") for l in self.lines if l.count(",") == 2]
        rows = rows[len(rows) // 3:]                          # the settled part
        pw = sorted(float(r[0]) for r in rows)
        ck = sorted(float(r[1]) for r in rows)
        return {"power_w_median": pw[len(pw) // 2] if pw else None, "sm_mhz_median": ck[len(ck) // 2] if ck else None,
                "power_cap_active_share": sum("Active" in r[2] for r in rows) / max(1, len(rows)), "samples": len(rows)}


dev = torch.device("cuda:0")
N, F = 65536, 2478461
lib = C.CDLL(os.path.join(HERE, "_build", "libgather_peak.so"))
lib.gather_mix.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
T = synth.make_motion_library(11313, seed=0, device=dev)
mlib = MotionLibSMPL.from_tables(T, device=dev)
frames, aux = mlib._packed if hasattr(mlib, "_packed") else None, None
frames = torch.empty(F * 312, device=dev).normal_()
aux = torch.empty(F * 48, device=dev).normal_()
sets = []
for s in range(4):
    g = torch.Generator(device="cpu").manual_seed(s)
    f0 = torch.randint(0, F - 1, (N,), generator=g).to(dev)
    sets.append((torch.randn(N * 312, device=dev), f0, f0 + 1, torch.empty(N * 1868, device=dev)))
sms = torch.cuda.get_device_properties(0).multi_processor_count
out = {}
smp = Sampler()


def timed(name, fn, seconds=3.0):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    smp.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n, t0 = 0, time.time()
    while time.time() - t0 < seconds:
        for i in range(200):
            fn(n + i)
        n += 200
        torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    r = smp.stop()
    r["ms_per_launch"] = a.elapsed_time(b) / n
    out[name] = r


def mover(i):
    sim, f0, f1, o = sets[i % 4]
    lib.gather_mix(sim.data_ptr(), frames.data_ptr(), aux.data_ptr(), f0.data_ptr(), f1.data_ptr(), N, o.data_ptr(), 2, sms * 8,
                   torch.cuda.current_stream().cuda_stream)


timed("traffic_only_kernel", mover)
del frames, aux, sets
rms = RunningNorm(934).to(dev)
fs = FusedStep(mlib, N, StepConfig(), rms=rms, normalize=True, accumulate_moments=True, defer_moments=True, metrics=True)
ins = [synth.make_env_state(T, N, seed=1 + s) for s in range(4)]
outs = [{"obs": torch.empty(N, 934, device=dev), "obs_norm": torch.empty(N, 934, device=dev)} for _ in range(4)]
keys = ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")


def step(i):
    fs(*[ins[i % 4][k] for k in keys], out=outs[i % 4])


timed("phc_step_fused", step)
print(json.dumps(out, indent=1))
