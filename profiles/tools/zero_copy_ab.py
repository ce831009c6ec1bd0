"""Experiment: the fused step reading its per-env inputs straight from pinned (UVA-mapped) host memory, no staging copies."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from puffer_phc_b200 import synth, _ffi
from puffer_phc_b200.fused_step import FusedStep, StepConfig
from puffer_phc_b200.motion_lib import MotionLibSMPL
from puffer_phc_b200.policies.running_norm import RunningNorm
dev = torch.device("cuda:0")
N = 65536
T = synth.make_motion_library(11313, seed=0, device=dev)
lib = MotionLibSMPL.from_tables(T, device=dev)
rms = RunningNorm(934).to(dev)
fs = FusedStep(lib, N, StepConfig(), rms=rms, normalize=True, accumulate_moments=True, defer_moments=True)
S = [synth.make_env_state(T, N, seed=1 + s) for s in range(2)]
hin = [{k: v.cpu().pin_memory() for k, v in s.items()} for s in S]
keys = ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")
want = {k: v.clone() for k, v in fs(*[S[0][k] for k in keys]).items()}
_ffi.require_cuda = lambda *a: None
import puffer_phc_b200.fused_step as F
F._ffi.require_cuda = lambda *a: None
got = fs(*[hin[0][k] for k in keys])
torch.cuda.synchronize()
ok = all(torch.equal(got[k], want[k]) for k in want)
for i in range(3): fs(*[hin[i % 2][k] for k in keys])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(16): fs(*[hin[i % 2][k] for k in keys])
b.record(); torch.cuda.synchronize()
print(json.dumps({"equal": ok, "zero_copy_ms": a.elapsed_time(b) / 16}))
