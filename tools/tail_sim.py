#!/usr/bin/env python3
"""Lane-efficiency model of the persistent ensemble kernel's tail (CPU only; attempt counts from the oracle).
Every lane pulls trajectories from a counter; a warp issues instructions until its last lane is done.
Compares: natural order, and "drain + sorted relaunch" (a warp hands its live lanes to a second launch as soon as one
of its lanes finds the counter exhausted; the second launch packs the hand-offs by progress t).
usage: tail_sim.py [log2 n] [lanes]"""
import heapq, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << lg
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else (94720 * n) >> 20
lanes -= lanes % 32
rng = np.random.default_rng(20261019)
y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
O.use_all_cores()
r = O.solve_ensemble("ode45", "lorenz", y0, [0.0, 10.0], [10.0, 28.0, 8.0 / 3.0])
att = (r["n_accept"] + r["n_reject"]).astype(np.int64)
print("n=%d lanes=%d attempts/traj mean %.1f min %d max %d" % (n, lanes, att.mean(), att.min(), att.max()))
# event simulation: lane i free at time f[i]
f = [(0, i) for i in range(lanes)]
heapq.heapify(f)
start = np.zeros(lanes, dtype=np.int64); cur = -np.ones(lanes, dtype=np.int64); end = np.zeros(lanes, dtype=np.int64)
for j in range(n):
    t, i = heapq.heappop(f)
    start[i] = t; cur[i] = j; end[i] = t + att[j]
    heapq.heappush(f, (end[i], i))
W = lanes // 32
endw = end.reshape(W, 32)
useful = att.sum()
base = 32 * endw.max(1).sum()
print("natural: lane efficiency %.4f (warp-iterations x32 = %.4g, useful %.4g)" % (useful / base, base, useful))
# drain: warp w stops at the first time one of its lanes finds no work = min over lanes of end (all lanes hold their
# last trajectory once the counter is exhausted at T0 = start of the last fetch)
T0 = start.max()
stopw = np.maximum(endw.min(1), T0)  # (a lane ending before T0 still got work)
A_iters = 32 * stopw.sum()
rem = np.maximum(endw - stopw[:, None], 0).reshape(-1)  # remaining attempts of every live lane at hand-off
live = rem > 0
# progress proxy the kernel can see: fraction of the time span done ~ attempts done / attempts total (noisy in reality)
attl = att[cur]
prog = 1.0 - rem / attl
noise = rng.normal(0, 0.05, prog.size)
for label, key in (("sorted by true remaining", rem.astype(float)), ("sorted by progress t (5% noise)", -(prog + noise) * attl.mean())):
    idx = np.argsort(-key[live])
    rr = rem[live][idx]
    pad = (-rr.size) % 32
    rr = np.concatenate([rr, np.zeros(pad, dtype=rr.dtype)]).reshape(-1, 32)
    B_iters = 32 * rr.max(1).sum()
    print("drain + relaunch, %s: lane efficiency %.4f (A %.4g + B %.4g), handed off %d lanes" % (label, useful / (A_iters + B_iters), A_iters, B_iters, live.sum()))
