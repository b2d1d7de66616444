"""Pinned host <-> device copy bandwidth of the box, one rank per GPU, all ranks copying at the same time.

    python scripts/pcie_bw.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_bw.py

Prints one JSON line per run (rank 0): per-rank and aggregate GB/s for H2D alone, D2H alone and both at once -- the ceiling
of bench.py's `e2e` figure when it is copy bound (bench.py measures the same thing per step as `e2e.copy_ceiling_ms`).
"""
import json
import os

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 512 * 1024 * 1024
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(h2d: bool, d2h: bool, reps: int = 5) -> float:
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ms = torch.tensor([best], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


res = {"n_gpus": world, "bytes_per_direction_per_rank": n}
for name, (a, b) in {"h2d": (True, False), "d2h": (False, True), "duplex": (True, True)}.items():
    ms = run(a, b)
    res[name] = {"ms": ms, "GBps_per_rank_per_direction": n / ms / 1e6, "GBps_aggregate": n * world * (2 if a and b else 1) / ms / 1e6}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
