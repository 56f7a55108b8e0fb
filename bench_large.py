"""bench.py --workload large: 1-D reaction-diffusion, method of lines, one system of
N = 2^26 unknowns per GPU, ode45 (Dormand-Prince), reference default tolerances,
t in [0, 20] (BASELINE.json configs[4]).

One benchmark "step" is one full integration (about 80 step attempts of the fused
attempt kernel + the control kernel).  value = whole-system step attempts per
second with u0 resident in HBM; e2e = the same through ode_b200_solve_large with
host buffers (512 MiB up, 512 MiB down per step).  Multi-GPU: the periodic grid
is split into slabs (weak scaling: N per GPU fixed), one all-gather per attempt.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FUSED_BYTES_PER_POINT = 32      # read y, k1; write ytrial, k7 (large-system kernel)
FIELD_BYTES_PER_POINT = 368     # one kernel per stage: 46 array passes for dopri5
STAGEWISE_BYTES_PER_POINT = 264  # SURVEY.md section 8d figure for one kernel per stage
FLOPS_PER_POINT = 143
# dram__bytes_read.sum + dram__bytes_write.sum of one attempt at N = 2^26 from the committed ncu capture
# (profiles/large_rd_r1.md: 1.075 GB + 1.034 GB) -- equal to the algorithmic 32 B/point, i.e. no re-reads
NCU_TRAFFIC_BYTES_PER_ATTEMPT_2_26 = 2.108e9


def u0_slab(lo, hi, n_global):
    i = np.arange(lo, hi, dtype=np.float64)
    return 0.5 + 0.4 * np.sin(2 * np.pi * 8 * i / n_global)


def cpu_baseline_large(n=1 << 22, attempts=12):
    """CPU restatement of ODE.jl (oracle, OpenMP over grid points) on a bounded sample: the first
    `attempts` step attempts of the same problem at N = 2^22."""
    from oracle import oracle as O
    y0 = u0_slab(0, n, n)
    threads = O.use_all_cores()
    O.solve("ode45", "reaction_diffusion", y0[:1 << 16], [0.0, 20.0], [1.0, 1.0], points="final", max_steps=2)  # warm-up
    t0 = time.perf_counter()
    s = O.solve("ode45", "reaction_diffusion", y0, [0.0, 20.0], [1.0, 1.0], points="final", max_steps=attempts)
    dt = time.perf_counter() - t0
    done = s.n_accept + s.n_reject
    return {"value": done * n / dt, "unit": "point-updates/s", "cores": threads, "kind": "port",
            "attempts_per_s_at_2^26": done * n / dt / (1 << 26),
            "sample": "first %d step attempts (hinit included) of the same problem at N = 2^22, %.2f s; C++ restatement "
                      "of ODE.jl (oracle/), OpenMP over grid points; Julia is not installed" % (done, dt)}


def run_large(args, rank, local_rank, world):
    import torch
    import ode_jl_b200 as m
    from bench import ClockSampler

    L = m.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_local = int(os.environ.get("ODE_B200_LARGE_N", 1 << 26))
    n_global = n_local * world
    lo, hi = m.shard_bounds(n_global, world, rank)
    field = getattr(args, "workload", "large") == "field"  # the same system through the stage-per-kernel path
    if field and world > 1:
        raise SystemExit("--workload field is a single-GPU path (a gather cannot be slab-decomposed)")
    F = m.DeviceRHS("reaction_diffusion_field" if field else "reaction_diffusion", 1.0, 1.0)
    bytes_per_point = FIELD_BYTES_PER_POINT if field else FUSED_BYTES_PER_POINT
    tspan = [0.0, 20.0]
    y0 = torch.from_numpy(u0_slab(lo, hi, n_global)).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"engine": None}

    def step():  # the slab session (device buffers) is created once and re-armed per integration
        y, ctl = m.solve_large_distributed("ode45", F, y0, n_global, tspan, engine=state["engine"], keep_engine=True)
        state["engine"] = ctl.pop("engine")
        return y, ctl

    steps = max(1, args.steps)
    for _ in range(max(3, args.warmup) if args.warmup else 1):
        y, ctl = step()
    barrier()
    attempts = ctl["attempts"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.ode_b200_launch_count(1)
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        y, ctl = step()  # state (4 x 512 MiB per attempt) is far larger than the 126 MB L2
    e1.record(stream)
    barrier()
    launches = int(L.ode_b200_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])

    e2e = None
    if world == 1:
        yh = y0.cpu().pin_memory().numpy()  # pinned host buffers for both legs
        outh = torch.empty(n_local, dtype=torch.float64).pin_memory().numpy()
        m.solve_large("ode45", F, yh, tspan, out=outh)
        t0 = time.perf_counter()
        r = m.solve_large("ode45", F, yh, tspan, out=outh)
        e2e_s = time.perf_counter() - t0
        assert np.array_equal(r["y"], y.cpu().numpy())
        e2e = {"value": (r["n_accept"] + r["n_reject"]) / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": yh.nbytes,
               "d2h_bytes_per_step": yh.nbytes, "ms_per_step": e2e_s * 1e3}
    else:
        # every rank: its slab from pinned host memory to the device, the distributed integration (public call
        # solve_large_distributed), the final slab back to pinned host memory; wall clock, max over ranks
        yh = y0.cpu().pin_memory()
        outh = torch.empty(n_local, dtype=torch.float64).pin_memory()

        def step_e2e():
            yd = yh.to(dev, non_blocking=True)
            yy, cc = m.solve_large_distributed("ode45", F, yd, n_global, tspan, engine=state["engine"], keep_engine=True)
            state["engine"] = cc.pop("engine")
            outh.copy_(yy, non_blocking=True)
            torch.cuda.synchronize()
            return cc

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        cc = step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
        assert torch.equal(outh, y.cpu())
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
        e2e = {"value": cc["attempts"] / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": yh.numel() * 8 * world,
               "d2h_bytes_per_step": yh.numel() * 8 * world, "ms_per_step": e2e_s * 1e3}
    if rank == 0:
        per_attempt_ms = ms / attempts
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        dmix = L.ode_b200_fp64_peak(local_rank, 1)
        gbs = bytes_per_point * n_local / (per_attempt_ms * 1e-3) / 1e9
        line = {"metric": "rk_step_attempts_per_s", "value": attempts / (ms * 1e-3), "unit": "steps/s",
                "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "reaction_diffusion_ode45_dp_field" if field else "reaction_diffusion_ode45_dp_large", "n_per_gpu": n_local, "n_global": n_global,
                           "tspan": tspan, "reltol": 1e-5, "abstol": 1e-8, "attempts_per_step": attempts,
                           "accepted": ctl["n_accept"], "rejected": ctl["n_reject"],
                           "ms_per_attempt": per_attempt_ms, "point_updates_per_s": attempts * n_global / (ms * 1e-3),
                           "l2": "working set 2 GiB per attempt >> 126 MB L2",
                           "parallelism": "slab x%d, one all-gather of %d doubles per attempt" % (world, m._lib.XCHG_DOUBLES)},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                             "traffic": NCU_TRAFFIC_BYTES_PER_ATTEMPT_2_26 if (n_local == (1 << 26) and not field) else None,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                             "algorithmic_bytes_per_point": bytes_per_point,
                             "stagewise_equivalent": {"bytes_per_point": STAGEWISE_BYTES_PER_POINT,
                                                      "gbs": STAGEWISE_BYTES_PER_POINT * n_local / (per_attempt_ms * 1e-3) / 1e9,
                                                      "frac": STAGEWISE_BYTES_PER_POINT * n_local / (per_attempt_ms * 1e-3) / 1e9 / hbm},
                             "fp64_issue_frac": FLOPS_PER_POINT * n_local / (per_attempt_ms * 1e-3) / dmix,
                             "note": ("general gather right-hand side: one kernel per stage, every stage argument "
                                      "materialised; HBM bound") if field else
                                     ("all stages of an attempt are fused in one kernel, so HBM traffic is 32 B/point "
                                      "instead of the 264 B/point of one kernel per stage; the kernel is FP64-issue bound")}}
        if world == 1 and not getattr(args, "no_cpu_baseline", False):
            line["cpu_baseline"] = cpu_baseline_large()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
