import sys, time, json, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from puffer_phc_b200 import synth
from puffer_phc_b200.c_gae import compute_gae_cuda
from puffer_phc_b200.fused_step import FusedStep, StepConfig
from puffer_phc_b200.motion_lib import MotionLibSMPL
from puffer_phc_b200.policies.running_norm import RunningNorm
dev='cuda:0'
T = synth.make_motion_library(11313, seed=0, device=dev)
lib = MotionLibSMPL.from_tables(T, device=dev)
for N in (1024, 4096, 16384, 65536):
    rms = RunningNorm(934).to(dev)
    fs = FusedStep(lib, N, StepConfig(), rms=rms, normalize=True, accumulate_moments=True)
    S = synth.make_env_state(T, N, seed=1)
    roll = synth.make_rollout(max(N // 32, 1), 32, seed=2, device=dev); adv = torch.empty_like(roll["rewards"])
    args = [S[k] for k in ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")]
    extra = lambda: compute_gae_cuda(roll["dones"], roll["values"], roll["rewards"], 0.98, 0.2, out=adv)
    def eager():
        fs(*args); extra()
    for name, fn in (("eager", eager),):
        for _ in range(20): fn()
        torch.cuda.synchronize(); a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        K = 500; a.record()
        for _ in range(K): fn()
        b.record(); torch.cuda.synchronize()
        print(N, name, f"{a.elapsed_time(b)/K*1e3:.1f} us/step  {N*K/(a.elapsed_time(b)*1e-3)/1e6:.1f} M env-steps/s")
    g, _ = fs.capture(*args, extra=extra)
    for _ in range(20): g.replay()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    K = 500; a.record()
    for _ in range(K): g.replay()
    b.record(); torch.cuda.synchronize()
    print(N, "graph", f"{a.elapsed_time(b)/K*1e3:.1f} us/step  {N*K/(a.elapsed_time(b)*1e-3)/1e6:.1f} M env-steps/s")
