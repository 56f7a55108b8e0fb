"""The large-system half of bench.py: 1-D reaction-diffusion, method of lines, ode45 (Dormand-Prince), reference
default tolerances, t in [0, 20] (BASELINE.json configs[4]; N = 2^26 per GPU).

One benchmark "step" is one full integration (about 80 step attempts of the fused attempt kernel, controller
included).  value = whole-system step attempts per second with u0 resident in HBM; e2e = the same through the
host-buffer entry point (pinned host u0 up, final state down into the caller's / the node-shared buffer).
Multi-GPU (one process per GPU): the periodic grid is split into slabs; per attempt the slabs exchange one record
of ODE_B200_XCHG_DOUBLES doubles -- written by the attempt kernel itself into every peer's HBM over NVLink (CUDA IPC
mappings), or, as a fallback, by an NCCL all-gather.  Before anything is timed the decomposed path is checked
against the single-slab bits on every rank.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FUSED_BYTES_PER_POINT = 32      # read y, k1; write ytrial, k7 (large-system kernel)
FIELD_BYTES_PER_POINT = 368     # one kernel per stage: 46 array passes for dopri5
FLOPS_PER_POINT = 143
TSPAN = [0.0, 20.0]
CSRC = os.path.join(ROOT, "ode.jl_b200", "csrc")
INC = os.path.join(ROOT, "include")
# the files a kernel is compiled from: its measured DRAM traffic (profiles/ncu_traffic.json) is only quoted while
# their hash is the one the capture was taken with
KERNEL_SOURCES = {
    "large": ["large.cuh", "common.cuh", "rhs.cuh", "tableaus.cuh"],
    "lorenz": ["ensemble_adapt.cuh", "common.cuh", "rhs.cuh", "tableaus.cuh"],
    "kepler": ["ensemble_adapt.cuh", "common.cuh", "rhs.cuh", "tableaus.cuh"],
    "robertson": ["ensemble_ode23s.cuh", "common.cuh", "rhs.cuh"],
    "field": ["field.cuh", "large.cuh", "common.cuh", "rhs.cuh", "tableaus.cuh"],
}


def source_sha(key):
    h = hashlib.sha256()
    for f in KERNEL_SOURCES[key]:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    for f in ("ode_b200_detmath.h", "ode_b200_detmath_tables.h"):
        h.update(open(os.path.join(INC, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(key, units):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the tracked ncu export, if it was captured with
    the kernel sources as they are now and for the same problem size; else None (stale)."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]
    except Exception:  # noqa: BLE001
        return None, "no tracked ncu capture"
    if rec.get("src_sha") != source_sha(key):
        return None, "stale: kernel sources changed since %s" % rec.get("capture")
    if int(rec.get("units", -1)) != int(units):
        return None, "capture was taken at another size (%s)" % rec.get("units")
    return float(rec["dram_bytes_per_launch"]), rec.get("capture")


def u0_slab(lo, hi, n_global):
    i = np.arange(lo, hi, dtype=np.float64)
    return 0.5 + 0.4 * np.sin(2 * np.pi * 8 * i / n_global)


def cpu_baseline_large(n=1 << 22, attempts=12):
    """CPU restatement of ODE.jl (oracle, OpenMP over grid points) on a bounded sample: the first
    `attempts` step attempts of the same problem at N = 2^22."""
    from oracle import oracle as O
    y0 = u0_slab(0, n, n)
    threads = O.use_all_cores()
    O.solve("ode45", "reaction_diffusion", y0[:1 << 16], TSPAN, [1.0, 1.0], points="final", max_steps=2)  # warm-up
    t0 = time.perf_counter()
    s = O.solve("ode45", "reaction_diffusion", y0, TSPAN, [1.0, 1.0], points="final", max_steps=attempts)
    dt = time.perf_counter() - t0
    done = s.n_accept + s.n_reject
    return {"value": done * n / dt / (1 << 26), "unit": "steps/s", "cores": threads, "kind": "port",
            "point_updates_per_s": done * n / dt,
            "sample": "first %d step attempts (hinit included) of the same problem at N = 2^22, %.2f s, scaled linearly to "
                      "attempts/s of a 2^26-point system; C++ restatement of ODE.jl (oracle/), OpenMP over grid points; "
                      "Julia is not installed" % (done, dt)}


class Ctx:
    """What every measurement needs: the package, the process group, this rank's device."""

    def __init__(self, rank, local_rank, world):
        import torch
        import ode_jl_b200 as m
        self.torch, self.m, self.L = torch, m, m.lib()
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def all_true(self, ok):
        if self.dist is None:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


def check_slabs_against_single_slab(ctx, n_check=1 << 20, exchange="auto"):
    """Real multi-rank parity, before anything is timed: the decomposed integration of an N = 2^20 system must give
    the single-slab attempt trace and this rank's slice of the single-slab final state, bit for bit, on every rank.
    Returns (ok, which exchange ran)."""
    torch, m = ctx.torch, ctx.m
    F = m.DeviceRHS("reaction_diffusion", 1.0, 1.0)
    y0 = u0_slab(0, n_check, n_check)
    single = m.solve_large("ode45", F, y0, TSPAN, trace=True, device=ctx.local_rank)
    lo, hi = m.shard_bounds(n_check, ctx.world, ctx.rank)
    y, ctl, tr = m.solve_large_distributed("ode45", F, torch.from_numpy(y0[lo:hi]).to(ctx.dev), n_check, TSPAN,
                                           trace=True, exchange=exchange)
    ok = (ctl["status"] == 0 and np.array_equal(tr.cpu().numpy(), single["trace"])
          and np.array_equal(y.cpu().numpy(), single["y"][lo:hi])
          and (ctl["n_accept"], ctl["n_reject"]) == (single["n_accept"], single["n_reject"]))
    m.lib().ode_b200_trim()
    return ctx.all_true(ok), ctl.get("exchange")


def measure_large(ctx, n_global, steps, warmup, *, field=False, exchange="auto", with_e2e=True, with_cpu=True,
                  scaling="weak", peaks=None):
    """Time `steps` full integrations of one N = n_global system split over ctx.world ranks.  Every rank calls this;
    the returned record is complete on rank 0."""
    torch, m, L = ctx.torch, ctx.m, ctx.L
    from bench import ClockSampler
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    lo, hi = m.shard_bounds(n_global, world, rank)
    n_local = hi - lo
    if field and world > 1:
        raise SystemExit("--workload field is a single-GPU path (a gather cannot be slab-decomposed)")
    F = m.DeviceRHS("reaction_diffusion_field" if field else "reaction_diffusion", 1.0, 1.0)
    bytes_per_point = FIELD_BYTES_PER_POINT if field else FUSED_BYTES_PER_POINT
    y0 = torch.from_numpy(u0_slab(lo, hi, n_global)).to(dev)
    state = {"engine": None, "exchange": None}

    def step(yd=y0):  # the slab session (device buffers, peer mappings) is created once and re-armed per integration
        y, ctl = m.solve_large_distributed("ode45", F, yd, n_global, TSPAN, engine=state["engine"], keep_engine=True,
                                           exchange=exchange)
        state["engine"] = ctl.pop("engine")
        state["exchange"] = ctl.get("exchange")
        return y, ctl

    steps = max(1, steps)
    for _ in range(max(3, warmup)):
        y, ctl = step()
    ctx.barrier()
    attempts = ctl["attempts"]
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    L.ode_b200_launch_count(1)
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = 0.0
    ctx.barrier()
    e0.record(stream)
    for _ in range(steps):
        y, ctl = step()  # the state (4 x 8 B x n_local per attempt) is far larger than the 126 MB L2 at the bench sizes
        kernel_ms += ctl.get("kernel_ms", 0.0)
    e1.record(stream)
    ctx.barrier()
    launches = int(L.ode_b200_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    # device time of the integration loops: the library's own CUDA events on the slab's stream (hinit .. last batch),
    # max over ranks; the torch events around the calls bound it from above (they include the per-call host work)
    ms_events = e0.elapsed_time(e1) / steps
    ms_lib = kernel_ms / steps if kernel_ms > 0 else ms_events
    ms_lib, ms = ctx.max_over_ranks(ms_lib, ms_events)  # `value` is quoted on the bracket around the K steps

    e2e = None
    if with_e2e:
        # every rank: its slab from pinned host memory to the device, the integration through the public call, the
        # final slab back into its slice of ONE host buffer shared by the ranks of the node (the final gather);
        # wall clock around the whole thing, max over ranks
        yh = y0.cpu().pin_memory()
        if world == 1:
            outh = torch.empty(n_local, dtype=torch.float64).pin_memory().numpy()
            yhn = yh.numpy()
            m.solve_large("ode45", F, yhn, TSPAN, out=outh, device=ctx.local_rank)
            t0 = time.perf_counter()
            r = m.solve_large("ode45", F, yhn, TSPAN, out=outh, device=ctx.local_rank)
            e2e_s = time.perf_counter() - t0
            assert np.array_equal(r["y"], y.cpu().numpy())
            att2 = r["n_accept"] + r["n_reject"]
            gather = "none (one rank)"
        else:
            from ode_jl_b200.sharding import SharedHostBlock
            blk = SharedHostBlock(n_global * 8, None)
            full = blk.view(np.float64, (n_global,))
            blk.pin(full[lo:hi])
            mine = torch.from_numpy(full[lo:hi])

            def step_e2e():
                yd = yh.to(dev, non_blocking=True)
                yy, cc = step(yd)
                mine.copy_(yy, non_blocking=True)
                torch.cuda.synchronize()
                ctx.dist.barrier()  # all slices are in the shared buffer: the gather is complete
                return cc

            step_e2e()
            ctx.barrier()
            t0 = time.perf_counter()
            cc = step_e2e()
            e2e_s = time.perf_counter() - t0
            ok = np.array_equal(full[lo:hi], y.cpu().numpy())
            if rank == 0 and n_global <= (1 << 27):  # rank 0 holds the whole final state without any further copy
                nxt = m.shard_bounds(n_global, world, 1)
                ok = ok and bool(np.all(np.isfinite(full[nxt[0]:nxt[0] + 1024])))
            assert ctx.all_true(ok), "e2e arm disagrees with the device-resident arm"
            att2 = cc["attempts"]
            gather = "each rank's D2H lands in its slice of one node-shared pinned host buffer"
            del mine, full
            blk.close()
        (e2e_s,) = ctx.max_over_ranks(e2e_s)
        e2e = {"value": att2 / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": n_global * 8,
               "d2h_bytes_per_step": n_global * 8, "ms_per_step": e2e_s * 1e3, "final_gather": gather}
    rec = None
    if rank == 0:
        per_attempt_ms = ms_lib / attempts  # the kernels alone: what the roofline is quoted on
        hbm = float((peaks or {}).get("hbm_gbs", 6650.0))
        dmix = L.ode_b200_fp64_peak(ctx.local_rank, 1)
        gbs = bytes_per_point * n_local / (per_attempt_ms * 1e-3) / 1e9
        traffic, tsrc = measured_traffic("field" if field else "large", n_local)
        rec = {"metric": "rk_step_attempts_per_s", "value": attempts / (ms * 1e-3), "unit": "steps/s",
               "n_gpus": world, "steps": steps, "warmup": max(3, warmup), "ms_per_step": ms, "higher_is_better": True,
               "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "reaction_diffusion_ode45_dp_field" if field else "reaction_diffusion_ode45_dp_large",
                          "n_per_gpu": n_local, "n_global": n_global, "tspan": TSPAN, "reltol": 1e-5, "abstol": 1e-8,
                          "attempts_per_step": attempts, "accepted": ctl["n_accept"], "rejected": ctl["n_reject"],
                          "ms_per_attempt": per_attempt_ms, "ms_per_step_kernels_only": ms_lib,
                          "point_updates_per_s": attempts * n_global / (ms * 1e-3),
                          "l2": "working set %.0f MiB per attempt vs 126 MB L2" % (4 * 8 * n_local / 2 ** 20),
                          "exchange": state["exchange"],
                          "parallelism": ("slab x%d, one record of %d doubles per rank per attempt" % (world, m._lib.XCHG_DOUBLES))
                          if world > 1 else "one slab"},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
               "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                            "traffic": traffic, "traffic_source": tsrc,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                            "algorithmic_bytes_per_point": bytes_per_point,
                            "fp64_issue_frac": FLOPS_PER_POINT * n_local / (per_attempt_ms * 1e-3) / dmix,
                            "fp64_issue_peak_instr_per_s": dmix,
                            "note": ("general gather right-hand side: one kernel per stage, every stage argument "
                                     "materialised; HBM bound") if field else
                                    ("all stages of an attempt are fused into one kernel: HBM traffic is the algorithmic "
                                     "32 B/point (SURVEY's stage-per-kernel design moves 264 B/point), which makes the "
                                     "kernel FP64-issue bound -- fp64_issue_frac is the fraction that limits it, the "
                                     "HBM fraction is what is left over")}}
        if with_cpu and world == 1:
            rec["cpu_baseline"] = cpu_baseline_large()
    if state["engine"] is not None:
        if world > 1:
            ctx.dist.barrier()
        state["engine"].close()
    L.ode_b200_trim()
    return rec
