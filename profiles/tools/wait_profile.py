"""Where do the compute warps of the fused step wait?  Needs a -DST_PROFILE=1 build:
    PHC_NVCC_EXTRA=-DST_PROFILE=1 python -m puffer_phc_b200.build --force && python profiles/tools/wait_profile.py
Prints, per compute-warp role, the share of the loop's cycles spent in: cp.async landing + group barrier, the plan barrier,
the tile-release barrier, the buffers-free group barrier."""
import ctypes as C, os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from puffer_phc_b200 import synth, _ffi
from puffer_phc_b200.fused_step import FusedStep, StepConfig
from puffer_phc_b200.motion_lib import MotionLibSMPL
from puffer_phc_b200.policies.running_norm import RunningNorm
dev = torch.device("cuda:0"); N = 65536
T = synth.make_motion_library(11313, seed=0, device=dev)
lib = MotionLibSMPL.from_tables(T, device=dev)
rms = RunningNorm(934).to(dev)
fs = FusedStep(lib, N, StepConfig(), rms=rms, normalize=True, accumulate_moments=True, defer_moments=True)
S = [synth.make_env_state(T, N, seed=1 + s) for s in range(4)]
keys = ("body_state", "progress", "start_time", "start_offset", "motion_ids", "global_offset", "dof_force", "dof_vel")
for i in range(12): fs(*[S[i % 4][k] for k in keys])
torch.cuda.synchronize()
buf = np.zeros(160 * 32 * 5, dtype=np.uint64)
so = _ffi.load()
so.phc_debug_profile.argtypes = [C.c_void_p]
assert so.phc_debug_profile(buf.ctypes.data) == 0
cw = int(os.environ.get("ST_CWARPS", "18"))          # compute warps: 6 per 4 env slots (default 12 slots)
p = buf.reshape(160, 32, 5)[:148, :cw].astype(np.float64)
names = ["landing+group", "plan", "tile_release", "group_barriers", "loop"]
out = {}
for role, sl in (("roleA", slice(0, cw // 2)), ("roleB", slice(cw // 2, cw))):
    q = p[:, sl]
    out[role] = {n: round(float(q[..., k].sum() / q[..., 4].sum()), 4) for k, n in enumerate(names[:4])}
    out[role]["loop_cycles_mean"] = float(q[..., 4].mean())
print(json.dumps(out, indent=1))
