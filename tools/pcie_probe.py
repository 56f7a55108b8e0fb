#!/usr/bin/env python3
"""Host <-> device copy bandwidth with all ranks copying at once (pinned memory, 512 MiB per direction per rank):
the ceiling the end-to-end numbers of a multi-rank run live under.  torchrun --nproc-per-node N tools/pcie_probe.py"""
import os, time
import torch
import torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True)),
                 ("both", None)):
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory() if name == "both" else None
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda") if name == "both" else None
    s2 = torch.cuda.Stream() if name == "both" else None
    for it in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            if name == "both":
                d.copy_(h, non_blocking=True)
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
            else:
                fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    gbs = 4 * n * (2 if name == "both" else 1) / dt / 1e9
    t = torch.tensor([gbs], device="cuda")
    if world > 1:
        lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(t, op=dist.ReduceOp.SUM)
        res[name] = (float(t.item()), float(lo.item()))
    else:
        res[name] = (gbs, gbs)
if rank == 0:
    print("ranks %d: " % world + ", ".join("%s aggregate %.1f GB/s (slowest rank %.1f)" % (k, v[0], v[1]) for k, v in res.items()))
if world > 1:
    dist.destroy_process_group()
