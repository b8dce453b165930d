#!/usr/bin/env python
"""PCIe ceiling of the box: pinned H2D, D2H and both at once (what the *_host pipeline competes with).

Single process:   python tools/pcie_check.py
All ranks at once (what limits the end-to-end figure at N GPUs: the ranks share the host's memory system and PCIe
root complexes):  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
                  --master-port 29517 tools/pcie_check.py
Rank 0 prints one JSON line: per-rank GB/s (min / mean / max over ranks) and the aggregate, for H2D alone, D2H alone and
both directions at once, every rank copying at the same time between barriers."""
import json
import os
import time

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:   # the ranks run next to their GPU, like bench.py does
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
    if cpus:
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
except Exception:
    cpus = set()

n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(h2d, d2h, reps=10):
    barrier(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9


def gather(v):
    if world == 1:
        return [v]
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


run(True, True, 2)
res = {}
for name, (a, b) in (("h2d_alone", (True, False)), ("d2h_alone", (False, True)), ("both_each_direction", (True, True))):
    v = gather(run(a, b))
    res[name] = {"min": min(v), "mean": sum(v) / len(v), "max": max(v), "aggregate": sum(v)}
if rank == 0:
    print(json.dumps({"what": "pinned-memory copy bandwidth, GB/s per rank, all ranks copying at once", "ranks": world,
                      "bytes_per_copy": n, "cpu_affinity_rank0": len(cpus), **res}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
