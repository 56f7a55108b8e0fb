#!/usr/bin/env python3
"""Times the fused attempt kernel + control kernel alone (no polling, no allocation
in the timed region).  Usage: python tools/large_probe.py [log2 N] [attempts] [method]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ode_jl_b200 as m

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
na = int(sys.argv[2]) if len(sys.argv) > 2 else 20
method = sys.argv[3] if len(sys.argv) > 3 else "ode45"
n = 1 << lg
dev = torch.device("cuda", 0)
i = torch.arange(n, dtype=torch.float64, device=dev)
y0 = 0.5 + 0.4 * torch.sin(2 * np.pi * 8 * i / n)
F = m.DeviceRHS("reaction_diffusion", 1.0, 1.0)
st = torch.cuda.current_stream(dev)
eng = m.CudaSlabEngine(method, F, n, n, 0, 1, [0.0, 20.0], st.cuda_stream)
send = torch.zeros(m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev)
recv = torch.zeros(1, m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev)
eng.set_state(y0.data_ptr())
for ph in range(4):
    eng.init_phase(ph, send.data_ptr(), recv.data_ptr())
    recv.view(-1).copy_(send)
torch.cuda.synchronize()
print("after hinit:", eng.poll())
for rep in range(2):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(st)
    for _ in range(na):
        eng.attempt(send.data_ptr())
    e1.record(st)
    for _ in range(na):
        eng.attempt(send.data_ptr())
        recv.view(-1).copy_(send)
        eng.control(recv.data_ptr())
    e2.record(st)
    torch.cuda.synchronize()
    a, b = e0.elapsed_time(e1) / na, e1.elapsed_time(e2) / na
    print("N=2^%d %s: attempt kernel alone %.4f ms; attempt+exchange+control %.4f ms; %.1f GB/s (32 B/pt), %.1f GB/s stagewise-equivalent (264 B/pt)"
          % (lg, method, a, b, 32 * n / a / 1e6, 264 * n / a / 1e6))
print(eng.poll())
