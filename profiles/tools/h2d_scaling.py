"""What limits the end-to-end (host-buffer) path when several GPUs of one box pull their per-env inputs at the same time?

    torchrun --nproc-per-node N profiles/tools/h2d_scaling.py            (one rank per GPU)

Every rank copies a 120 MB pinned host buffer (the per-step input volume of 65536 envs) to its GPU, (a) alone, one rank after the
other, (b) all ranks at once, for three placements of the pinned buffer: wherever the allocating thread happens to run (what
FusedStep.step_host did in round 1), on the NUMA node the GPU hangs off (thread pinned to that node's cores BEFORE the allocation:
cudaHostAlloc takes its pages from the calling thread's node), and on the OTHER node.  Rank 0 prints the table and the box's
topology (nvidia-smi topo -m, NUMA nodes), so the limiter -- PCIe link, PCIe switch uplink, inter-socket link or host DRAM --
can be named from the numbers.
"""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from puffer_phc_b200 import hostmem  # noqa: E402

MB = 120


def bw(buf_h, buf_d, iters=10):
    s = torch.cuda.current_stream()
    for _ in range(2):
        buf_d.copy_(buf_h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(iters):
        buf_d.copy_(buf_h, non_blocking=True)
    b.record(s)
    torch.cuda.synchronize()
    return buf_h.numel() * iters / (a.elapsed_time(b) * 1e-3) / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    buf_d = torch.empty(MB << 20, dtype=torch.uint8, device=dev)
    node = hostmem.gpu_numa_node(local)
    nodes = hostmem.numa_nodes()
    res = {"rank": rank, "gpu_numa_node": node, "numa_nodes": sorted(nodes)}
    placements = {"default": None, "gpu_node": node}
    others = [n for n in nodes if n != node]
    if others:
        placements["other_node"] = others[0]
    for name, n in placements.items():
        with hostmem.on_numa_node(n):
            h = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
            h.fill_(1)
        alone = []
        for r in range(world):                       # one rank at a time
            if world > 1:
                dist.barrier()
            if r == rank:
                alone.append(bw(h, buf_d))
        if world > 1:
            dist.barrier()
        together = bw(h, buf_d, iters=20)            # all ranks at once
        res[name] = {"alone_GBps": alone[0], "concurrent_GBps": together}
        del h
    out = [None] * world
    if world > 1:
        dist.all_gather_object(out, res)
    else:
        out = [res]
    if rank == 0:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        lscpu = subprocess.run("lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout
        agg = {name: sum(o[name]["concurrent_GBps"] for o in out) for name in placements}
        print(json.dumps({"world": world, "mb_per_copy": MB, "ranks": out, "aggregate_concurrent_GBps": agg, "topo": topo, "lscpu": lscpu}, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
