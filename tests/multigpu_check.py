#!/usr/bin/env python3
"""(Not collected by pytest: needs torchrun and >= 2 GPUs.)  Checks the two sharded paths with one process per GPU
on real hardware against the CPU oracle:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py
 * ensembles: trajectory shards + the final gather (node-shared host buffer);
 * large system: slabs with the exchange over peer memory (CUDA IPC, kernels write into each other's HBM) AND with
   the NCCL all-gather per attempt; both must equal the single-process oracle bit for bit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import ode_jl_b200 as m
from oracle import oracle as O

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
# ensembles: shards + final gather
n = 20001
rng = np.random.default_rng(5)
y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
F = m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3)
got = m.solve_ensemble_distributed("ode45", F, y0, [0.0, 2.0])
idx = rng.choice(n, 512, replace=False)
ref = O.solve_ensemble("ode45", "lorenz", y0[idx], [0.0, 2.0], [10.0, 28.0, 8 / 3])
ok1 = all(np.array_equal(got[k][idx], ref[k]) for k in ("y_final", "n_accept", "n_reject", "status"))
# large system: slabs, against the single-process oracle
N = 30011
u0 = 0.5 + 0.4 * np.sin(2 * np.pi * 8 * np.arange(N) / N)
lo, hi = m.shard_bounds(N, world, rank)
o = O.solve("ode45", "reaction_diffusion", u0, [0.0, 20.0], [1.0, 1.0], trace=True, points="final")
oks = {}
for mode in ("peer", "nccl"):
    t0 = time.perf_counter()
    y, ctl, tr = m.solve_large_distributed("ode45", m.DeviceRHS("reaction_diffusion", 1.0, 1.0),
                                           torch.tensor(u0[lo:hi], device="cuda"), N, [0.0, 20.0], trace=True,
                                           exchange=mode)
    el = time.perf_counter() - t0
    ok = (ctl["exchange"] == mode and ctl["status"] == 0 and np.array_equal(y.cpu().numpy(), o.y[-1][lo:hi])
          and np.array_equal(tr.cpu().numpy(), o.trace))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    oks[mode] = bool(flag.item())
    if rank == 0:
        print("large slabs x%d exchange=%s bit-exact: %s | attempts %d | %.1f ms incl. session setup"
              % (world, mode, oks[mode], ctl["attempts"], el * 1e3))
if rank == 0:
    print("world", world, "ensemble shards+gather bit-exact:", ok1)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok1 and all(oks.values())) else 1)
