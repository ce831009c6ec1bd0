"""Time the device-side auto-reset (phc_auto_reset: scan + tail kernels) on the step's own outputs at 65536 envs.

    python profiles/tools/reset_bench.py [iters]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from puffer_phc_b200 import synth                                   # noqa: E402
from puffer_phc_b200.envs.reset import AutoReset, EnvTensors        # noqa: E402
from puffer_phc_b200.fused_step import FusedStep, StepConfig        # noqa: E402
from puffer_phc_b200.motion_lib import MotionLibSMPL                # noqa: E402
from puffer_phc_b200.policies.running_norm import RunningNorm       # noqa: E402

dev = torch.device("cuda:0")
N = 65536
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T = synth.make_motion_library(11313, seed=0, device=dev)
lib = MotionLibSMPL.from_tables(T, device=dev)
rms = RunningNorm(934).to(dev)
fs = FusedStep(lib, N, StepConfig(), rms=rms, normalize=True, accumulate_moments=True, defer_moments=True, metrics=True)
S = synth.make_env_state(T, N, seed=1)
keys = ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")
pristine = {k: S[k].clone() for k in ("body_state", "progress", "start_time", "start_offset", "global_offset")}
env = EnvTensors(rigid_body_state=S["body_state"], humanoid_root_states=torch.empty(N, 13, device=dev), dof_pos=torch.empty(N, 69, device=dev),
                 dof_vel=S["dof_vel"], progress_buf=S["progress"], reset_buf=fs.reset_buf, terminate_buf=fs.terminate_buf,
                 global_offset=S["global_offset"], motion_start_times=S["start_time"], motion_start_times_offset=S["start_offset"],
                 sampled_motion_ids=S["motion_ids"], obs_buf=fs.obs_buf)
ar = AutoReset(env, lib, obs_norm=fs.obs_norm, rms=rms, fused=fs)
phase = torch.rand(N, device=dev)
a = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
b = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
for i in range(-3, iters):
    fs(*[S[k] for k in keys])
    if i >= 0:
        a[i].record()
    ar(fs.rew_buf, fs.reward_raw, phase)
    if i >= 0:
        b[i].record()
    for k, v in pristine.items():
        S[k].copy_(v)
torch.cuda.synchronize()
us = sum(x.elapsed_time(y) for x, y in zip(a, b)) / iters * 1e3
print(json.dumps({"envs": N, "flagged": int(ar.reset_count), "us_auto_reset": us}))
