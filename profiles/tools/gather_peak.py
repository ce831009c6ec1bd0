"""Ceiling of the fused step's access pattern (see gather_peak.cu): GB/s of a kernel that only moves the same bytes.

    python profiles/tools/gather_peak.py          (needs profiles/tools/_build/libgather_peak.so, built by the nvcc line in the .cu)
"""
import ctypes as C
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "_build", "libgather_peak.so"))
lib.gather_mix.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
dev = torch.device("cuda:0")
N, F = 65536, 2478461                                    # envs, frames of the 11313-clip library (3.1 GB of records)
frames = torch.empty(F * 312, device=dev).normal_()
aux = torch.empty(F * 48, device=dev).normal_()
sets = []
for s in range(4):
    g = torch.Generator(device="cpu").manual_seed(s)
    f0 = torch.randint(0, F - 1, (N,), generator=g).to(dev)
    sets.append((torch.randn(N * 312, device=dev), f0, f0 + 1, torch.empty(N * 1868, device=dev)))
out = {}
sms = torch.cuda.get_device_properties(0).multi_processor_count
for fpe, label in ((1, "one frame per query (blend == 0 fast path: 84 % of the step's queries)"), (2, "two frames per query")):
    for blocks_per_sm in (4, 8):
        def run(i):
            sim, f0, f1, o = sets[i % 4]
            lib.gather_mix(sim.data_ptr(), frames.data_ptr(), aux.data_ptr(), f0.data_ptr(), f1.data_ptr(), N, o.data_ptr(), fpe,
                           sms * blocks_per_sm, torch.cuda.current_stream().cuda_stream)
        for i in range(10):
            run(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(200):
            run(i)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 200
        rd = 1248 + 16 + fpe * 1248 + 2 * 192 if fpe == 2 else 1248 + 16 + 1248 + 192 + 192
        by = (rd + 7472) * N
        out[f"frames{fpe}_blocks{blocks_per_sm}"] = {"ms": ms, "GBps": by / ms / 1e6, "bytes_per_env": rd + 7472, "what": label}
print(json.dumps(out, indent=1))
