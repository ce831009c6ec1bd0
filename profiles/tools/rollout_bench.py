"""Time the device-resident rollout buffer (row f2): store x T, then sort_training_data + compute_advantages, at [T=32, N] sizes.

    python profiles/tools/rollout_bench.py
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from puffer_phc_b200.rollout import RolloutBuffer  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for N in (4096, 65536):
    T = 32
    g = torch.Generator(device="cpu").manual_seed(3)
    vals = torch.randn(T + 4, N, generator=g).to(dev)
    rews = torch.rand(T + 4, N, generator=g).to(dev)
    dones = (torch.rand(T + 4, N, generator=g) < 0.01).float().to(dev)
    for name, p_trunc in (("regular", 0.0), ("ragged_1pct_truncated", 0.01)):
        trunc = (torch.rand(T + 4, N, generator=g) < p_trunc).to(dev)
        buf = RolloutBuffer(N, N * T, device=dev)

        def rollout():
            buf.reset()
            for t in range(T + 4):
                buf.store(vals[t], rews[t], dones[t], trunc[t].float(), ~trunc[t])
                if p_trunc == 0.0 and t == T - 1:
                    break
        def post():
            buf.sort_training_data()
            return buf.compute_advantages(0.98, 0.2)
        for _ in range(3):
            rollout(); post()
        torch.cuda.synchronize()
        ts, tp = [], []
        for _ in range(10):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            rollout()
            torch.cuda.synchronize(); t1 = time.perf_counter()
            post()
            torch.cuda.synchronize(); t2 = time.perf_counter()
            ts.append(t1 - t0); tp.append(t2 - t1)
        out[f"{N}x{T}_{name}"] = {"store_all_steps_us": 1e6 * sorted(ts)[len(ts) // 2], "sort_plus_gae_us": 1e6 * sorted(tp)[len(tp) // 2],
                                  "rows": int(buf.idxs.numel())}
print(json.dumps(out, indent=1))
