"""A/B of FusedStep.step_host chunk counts + the raw pinned H2D bandwidth of the box (one 82 MB tensor)."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from puffer_phc_b200 import synth
from puffer_phc_b200.fused_step import FusedStep, StepConfig
from puffer_phc_b200.motion_lib import MotionLibSMPL
from puffer_phc_b200.policies.running_norm import RunningNorm
dev = torch.device("cuda:0")
N = 65536
T = synth.make_motion_library(11313, seed=0, device=dev)
lib = MotionLibSMPL.from_tables(T, device=dev)
rms = RunningNorm(934).to(dev)
fs = FusedStep(lib, N, StepConfig(), rms=rms, normalize=True, accumulate_moments=True, defer_moments=True)
hin = [{k: v.cpu().pin_memory() for k, v in synth.make_env_state(T, N, seed=1 + s).items()} for s in range(2)]
out = {}
big = hin[0]["body_state"]; dbig = torch.empty_like(big, device=dev)
for _ in range(3): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): dbig.copy_(big, non_blocking=True)
b.record(); torch.cuda.synchronize()
out["raw_h2d_gbs"] = big.numel() * 4 * 10 / (a.elapsed_time(b) * 1e-3) / 1e9
for chunks, hs in ((1, 1), (2, 1), (4, 1), (2, 2), (4, 2), (4, 3), (8, 2)):
    for i in range(3): fs.step_host(hin[i % 2], chunks=chunks, h2d_streams=hs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(16): fs.step_host(hin[i % 2], chunks=chunks, h2d_streams=hs)
    torch.cuda.synchronize()
    out[f"chunks{chunks}_h2d{hs}_ms"] = (time.perf_counter() - t0) / 16 * 1e3
# two steps in flight (wait=False handles, double-buffered staging)
for chunks, hs in ((1, 1), (2, 1), (4, 1), (1, 2), (2, 2), (4, 2)):
    for i in range(3): fs.step_host(hin[i % 2], chunks=chunks, h2d_streams=hs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pending = None
    for i in range(16):
        h = fs.step_host(hin[i % 2], chunks=chunks, h2d_streams=hs, wait=False)
        if pending is not None: pending.result()
        pending = h
    pending.result()
    torch.cuda.synchronize()
    out[f"inflight2_chunks{chunks}_h2d{hs}_ms"] = (time.perf_counter() - t0) / 16 * 1e3
print(json.dumps(out))
