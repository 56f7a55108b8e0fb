#!/usr/bin/env python3
"""(Not collected by pytest: needs torchrun and >= 2 GPUs.)  Checks the two sharded paths on real hardware
(NCCL) against the CPU oracle.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py"""
import os, sys
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
# large system: slabs + all-gather per attempt, against the single-process oracle
N = 30011
u0 = 0.5 + 0.4 * np.sin(2 * np.pi * 8 * np.arange(N) / N)
lo, hi = m.shard_bounds(N, world, rank)
y, ctl, tr = m.solve_large_distributed("ode45", m.DeviceRHS("reaction_diffusion", 1.0, 1.0),
                                       torch.tensor(u0[lo:hi], device="cuda"), N, [0.0, 20.0], trace=True)
parts = [torch.zeros(m.shard_bounds(N, world, r)[1] - m.shard_bounds(N, world, r)[0], dtype=torch.float64, device="cuda")
         for r in range(world)]
pad = max(p.numel() for p in parts)
buf = torch.zeros(pad, dtype=torch.float64, device="cuda")
buf[:y.numel()] = y
allb = [torch.zeros(pad, dtype=torch.float64, device="cuda") for _ in range(world)]
dist.all_gather(allb, buf)
yy = np.concatenate([allb[r][:parts[r].numel()].cpu().numpy() for r in range(world)])
o = O.solve("ode45", "reaction_diffusion", u0, [0.0, 20.0], [1.0, 1.0], trace=True, points="final")
ok2 = np.array_equal(yy, o.y[-1]) and np.array_equal(tr.cpu().numpy(), o.trace)
if rank == 0:
    print("world", world, "ensemble shards+gather bit-exact:", ok1, "| large slabs bit-exact:", ok2,
          "| attempts", ctl["attempts"])
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok1 and ok2) else 1)
