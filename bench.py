#!/usr/bin/env python3
"""bench.py -- headline benchmark: RK step attempts per second.

Headline workload (BASELINE.json configs[1]): Lorenz, ode45 (Dormand-Prince), 2^20
trajectories with randomised initial conditions, FP64, t in [0, 10], reference
default tolerances (reltol 1e-5, abstol 1e-8).  A "step" of this benchmark is
one pass of the hot path over the batch, i.e. the whole ensemble solve; the
metric counts RK step attempts (accepted + rejected, one attempt = one
rk_embedded_step! + stepsize_hw92!, /root/reference src/runge_kutta.jl:305-310).

The ONE JSON line of the default run (`--workload all`) is that headline record plus
  "large"  : the second north-star regime measured the same way -- reaction-diffusion, N = 2^26 per GPU,
             ode45, one fused kernel per step attempt (bench_large.py) -- with its own value, e2e,
             roofline and cpu_baseline; on N > 1 ranks the grid is cut into slabs that exchange one
             record per attempt over NVLink peer memory (checked against the single-slab bits first);
  "strong" : N > 1 only: the same two workloads at FIXED total size (2^20 trajectories, N = 2^26)
             split over the ranks, next to the weak-scaling numbers above.
`--workload lorenz | large | field` runs one of them alone.

  value : attempts/s with y0 already resident in HBM (ode_b200_solve_ensemble_dev
          on torch's current stream, CUDA events around every step)
  e2e   : the same through the host-buffer C-ABI call ode_b200_solve_ensemble
          (pinned host y0 -> H2D, kernel, D2H of every output) per step
  roofline   : algorithmic FP64 flops (297 per attempt) per
               second against the FP64 peak measured live with the library's
               DFMA / DADD+DMUL microbenchmark (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline : the CPU oracle (C++ restatement of ODE.jl, OpenMP over
               trajectories) on a bounded sample, on rank 0 at N=1 only

`--impl reference` times that CPU restatement instead (the reference itself is
Julia, which is not installed here or on the GPU box).
Multi-GPU (torchrun, one process per GPU): trajectories shard across ranks with no
data-path collective (weak scaling: 2^20 trajectories per GPU); the final gather is
every rank's D2H into its slice of one node-shared pinned host buffer, inside e2e.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LORENZ = (10.0, 28.0, 8.0 / 3.0)
FLOPS_PER_ATTEMPT = 297  # dopri5 / Lorenz, SURVEY.md section 8d
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 148 SMs x 64 FP64 FMA lanes x 2 flops x 1.965 GHz = 37.2
SEED = 20261019


def lorenz_ics(n, seed=SEED):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_lorenz(sample, steps=1, warmup=0):
    """CPU restatement of ODE.jl (oracle, OpenMP over trajectories) on `sample` trajectories."""
    from oracle import oracle as O
    y0 = lorenz_ics(sample)
    threads = O.use_all_cores()
    best = None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = O.solve_ensemble("ode45", "lorenz", y0, [0.0, 10.0], list(LORENZ), threads=threads)
        dt = time.perf_counter() - t0
        att = int(r["n_accept"].sum() + r["n_reject"].sum())
        if it >= warmup:
            best = dt if best is None else best + dt
    per = best / steps
    return dict(value=att / per, unit="steps/s", cores=threads, kind="port",
                sample="%d of the 2^20 Lorenz/ode45 trajectories (same seed, first rows), %d attempts, %.2f s per pass; "
                       "C++ restatement of ODE.jl (oracle/), OpenMP dynamic schedule; Julia is not installed"
                       % (sample, att, per)), per, att


def bind_to_gpu_numa(local_rank):
    """Several ranks on one host: run this rank (and so first-touch its pinned buffers) on the NUMA node its GPU
    hangs off.  Returns a short description for the bench record; never fails."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]  # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa_node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, cpus)
        return "bound to numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:  # noqa: BLE001
        return "not bound (%s)" % type(e).__name__


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port."""
    if rank != 0:
        return
    if args.workload in ("large", "field"):
        from bench_large import cpu_baseline_large
        cb = cpu_baseline_large()
        v = cb["value"]
        print(json.dumps({"impl": "reference", "metric": "rk_step_attempts_per_s", "value": v, "unit": "steps/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic",
                          "config": {"workload": "reaction_diffusion_ode45_dp_large", "n_per_gpu": 1 << 26,
                                     "note": "attempts/s of a 2^26-point system extrapolated linearly from 2^22 points"},
                          "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return
    sample = args.cpu_sample
    cb, per, att = cpu_baseline_lorenz(sample, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "rk_step_attempts_per_s", "value": cb["value"], "unit": "steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "lorenz_ode45_dp_ensemble", "n_traj_per_gpu": 1 << 20, "tspan": [0.0, 10.0],
                       "reltol": 1e-5, "abstol": 1e-8, "sampled_trajectories": sample},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload == "all":  # the second regime's CPU number rides along, like the `large` record of our arm
        from bench_large import cpu_baseline_large
        cbl = cpu_baseline_large()
        line["large"] = {"impl": "reference", "metric": "rk_step_attempts_per_s", "value": cbl["value"], "unit": "steps/s",
                         "higher_is_better": True, "config": {"workload": "reaction_diffusion_ode45_dp_large", "n_per_gpu": 1 << 26},
                         "cpu_baseline": cbl,
                         "e2e": {"value": cbl["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def measure_lorenz(ctx, n, y0_np, steps, warmup, *, scaling, peaks, with_cpu, cpu_sample, strict_fp=True):
    """`steps` passes of the ode45 hot path over this rank's n Lorenz trajectories; record complete on rank 0."""
    import ctypes as C
    torch, m, L = ctx.torch, ctx.m, ctx.L
    from ode_jl_b200 import _lib
    from bench_large import measured_traffic
    rank, world, dev, local_rank = ctx.rank, ctx.world, ctx.dev, ctx.local_rank
    dist = ctx.dist
    dof = 3
    y0_host = torch.from_numpy(np.ascontiguousarray(y0_np)).pin_memory()
    F = m.DeviceRHS("lorenz", *LORENZ)
    tspan = np.array([0.0, 10.0])
    opts = m.make_opts("ode45", F, tspan, points="final", device=local_rank, strict_fp=strict_fp)

    # ---- device-resident arm ------------------------------------------------------
    d_y0 = y0_host.to(dev)
    d_par = torch.tensor(LORENZ, dtype=torch.float64, device=dev)
    d_ts = torch.tensor(tspan, dtype=torch.float64, device=dev)
    d_tf = torch.empty(n, dtype=torch.float64, device=dev)
    d_yf = torch.empty(n, dof, dtype=torch.float64, device=dev)
    d_na = torch.empty(n, dtype=torch.int32, device=dev)
    d_nr = torch.empty(n, dtype=torch.int32, device=dev)
    d_nf = torch.empty(n, dtype=torch.int32, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    io = _lib.EnsembleIO()
    io.n_traj, io.n_tspan, io.trace_traj = n, 2, -1
    io.y0, io.params, io.tspan = d_y0.data_ptr(), d_par.data_ptr(), d_ts.data_ptr()
    io.t_final, io.y_final = d_tf.data_ptr(), d_yf.data_ptr()
    io.n_accept, io.n_reject, io.n_rhs, io.status = d_na.data_ptr(), d_nr.data_ptr(), d_nf.data_ptr(), d_st.data_ptr()
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_dev():
        _lib.check(L.ode_b200_solve_ensemble_dev(C.byref(opts), C.byref(io), C.c_void_p(stream.cuda_stream)))

    for _ in range(warmup):
        step_dev()
    ctx.barrier()
    attempts = int((d_na.sum() + d_nr.sum()).item())
    assert int(d_st.max().item()) == 0, "a trajectory did not finish"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.ode_b200_launch_count(1)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    for e0, e1 in evs:
        flush.fill_(1)  # L2 flush between timed iterations (not timed)
        e0.record(stream)
        step_dev()
        e1.record(stream)
    ctx.barrier()
    launches = int(L.ode_b200_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs) / steps

    # ---- end-to-end arm: host buffers through the C ABI ----------------------------
    # One rank: caller-owned pinned output buffers.  Several ranks: the outputs ARE the final gather -- every rank's
    # library call copies its results device -> host into its slice of one node-shared page-locked buffer, and a
    # barrier per step closes the gather (north star: "no collectives beyond the final gather").
    shared = None
    if world == 1:
        out = dict(t_final=torch.empty(n, dtype=torch.float64).pin_memory().numpy(),
                   y_final=torch.empty(n, dof, dtype=torch.float64).pin_memory().numpy(),
                   n_accept=torch.empty(n, dtype=torch.int32).pin_memory().numpy(),
                   n_reject=torch.empty(n, dtype=torch.int32).pin_memory().numpy(),
                   status=torch.empty(n, dtype=torch.int32).pin_memory().numpy())
    else:
        (n_max,) = ctx.max_over_ranks(n)
        assert int(n_max) == n, "equal shards expected"
        shared = m.SharedEnsembleResult(n * world, dof, None)
        out = shared.mine
    par_h = np.array(LORENZ)
    io2 = _lib.EnsembleIO()
    io2.n_traj, io2.n_tspan, io2.trace_traj = n, 2, -1
    io2.y0, io2.params, io2.tspan = y0_host.data_ptr(), par_h.ctypes.data, tspan.ctypes.data
    io2.t_final, io2.y_final = out["t_final"].ctypes.data, out["y_final"].ctypes.data
    io2.n_accept, io2.n_reject, io2.status = out["n_accept"].ctypes.data, out["n_reject"].ctypes.data, out["status"].ctypes.data
    h2d = y0_host.numel() * 8 + 3 * 8 + 2 * 8
    d2h = n * 8 + n * dof * 8 + 3 * n * 4
    # direct = the kernel writes the results into the pinned output buffers itself (host_output = 1, no copy-back
    # phase).  Default: on for one rank (11.05 vs 11.25 ms); several ranks per host take the progressive copy-back of
    # the staged path: the host cannot absorb eight kernels' small writes (8 x B200: 15.3 ms direct, 14.8 ms with a
    # copy phase after the kernel, 11.46 ms progressive).
    direct = world == 1
    if os.environ.get("ODE_B200_BENCH_DIRECT"):
        direct = os.environ["ODE_B200_BENCH_DIRECT"] != "0"
    opts_e2e = m.make_opts("ode45", F, tspan, points="final", device=local_rank, direct_output=direct, strict_fp=strict_fp)

    def step_e2e():
        _lib.check(L.ode_b200_solve_ensemble(C.byref(opts_e2e), C.byref(io2)))
        if dist is not None:  # every rank's slice is in the shared buffer
            if os.environ.get("ODE_B200_BENCH_BARRIER") == "nccl":
                dist.barrier()
            else:
                shared.block.host_barrier()

    for _ in range(2):
        step_e2e()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    ctx.barrier()
    e2e_s = (time.perf_counter() - t0) / steps
    L.ode_b200_last_kernel_ms.restype = C.c_double
    e2e_kernel_ms = float(L.ode_b200_last_kernel_ms())  # device span of the last call's kernels (events inside the library)
    att2 = int(out["n_accept"].sum() + out["n_reject"].sum())
    assert att2 == attempts
    assert np.array_equal(out["y_final"], d_yf.cpu().numpy()), "host-buffer and device-resident arms disagree"
    tot_attempts = attempts
    if world > 1:
        if rank == 0:  # the gathered ensemble is complete on rank 0 (and everywhere else)
            assert int(shared.arrays["n_accept"].sum() + shared.arrays["n_reject"].sum()) > attempts
            assert int(shared.arrays["status"].max()) == 0
        dev_ms, e2e_s = ctx.max_over_ranks(dev_ms, e2e_s)
        a = torch.tensor([attempts], dtype=torch.float64, device=dev)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        tot_attempts = int(a.item())
        shared.close()

    line = None
    if rank == 0:
        value = tot_attempts / (dev_ms * 1e-3)
        e2e_v = tot_attempts / e2e_s
        # roofline of the dominant kernel (this rank's launch): algorithmic flops / kernel time
        dfma = L.ode_b200_fp64_peak(local_rank, 0)   # FP64 FMA instr/s, measured now
        dmix = L.ode_b200_fp64_peak(local_rank, 1)   # DADD+DMUL instr/s, measured now
        achieved = attempts * FLOPS_PER_ATTEMPT / (dev_ms * 1e-3) / 1e12
        peak = 2.0 * dfma / 1e12
        traffic, tsrc = measured_traffic("lorenz", n)
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "frac_of_issue_peak": achieved / (max(dfma, dmix) / 1e12),
                    "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
                    "traffic": traffic, "traffic_source": tsrc,
                    "peak_source": "measured live: ode_b200_fp64_peak DFMA chains x2 flops (no FP64 entry in "
                                   "MEASURED_PEAKS.json; ncu of the microbenchmark: profiles/); issue peak = "
                                   "max(DFMA, DADD+DMUL) instr/s = %.3g; nominal = 148 SM x 64 lanes x 2 x 1.965 GHz" % max(dfma, dmix),
                    "note": ("strict no-FMA arithmetic (reference parity) caps flops at 50% of the FMA peak; "
                             "frac_of_issue_peak counts algorithmic FP64 ops against FP64 instruction slots") if strict_fp else
                            "fp_mode = FMA: contracted multiply-adds, NOT the parity path (flops counted as in strict mode)"}
        line = {"metric": "rk_step_attempts_per_s", "value": value, "unit": "steps/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": dev_ms, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "lorenz_ode45_dp_ensemble", "n_traj_per_gpu": n, "tspan": [0.0, 10.0],
                           "reltol": 1e-5, "abstol": 1e-8, "attempts_per_step_per_gpu": attempts,
                           "fp_mode": "strict (no FMA, reference parity)" if strict_fp else "fma (opt-in, not bit-exact)",
                           "l2": "256 MiB flush write between timed iterations", "parallelism": "traj-shard x%d" % world},
                "e2e": {"output_path": ("kernel writes pinned host buffers" if direct else
                                        "results staged in HBM, copied back chunk by chunk while the kernel runs") +
                                       ("" if world == 1 else "; final gather = every rank's results land in its slice "
                                                              "of one node-shared pinned host buffer, barrier per step"),
                        "value": e2e_v, "unit": "steps/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "ms_per_step": e2e_s * 1e3, "kernels_ms_inside_last_call": e2e_kernel_ms},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if with_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_lorenz(cpu_sample)[0]
    del flush
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "lorenz", "large", "field"])
    ap.add_argument("--n-traj", type=int, default=1 << 20)
    ap.add_argument("--large-n", type=int, default=int(os.environ.get("ODE_B200_LARGE_N", 1 << 26)))
    ap.add_argument("--cpu-sample", type=int, default=1 << 18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"])
    ap.add_argument("--relaxed-fp", action="store_true", help="opt-in FMA-contracted kernels (not the parity path)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    numa = bind_to_gpu_numa(local_rank) if world > 1 else "one rank: not bound"
    import torch
    from bench_large import Ctx, measure_large, check_slabs_against_single_slab
    ctx = Ctx(rank, local_rank, world)
    if not torch.cuda.is_available() or ctx.L.ode_b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        ctx.dist.init_process_group("nccl", device_id=ctx.dev)
    peaks = None
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    with_cpu = not args.no_cpu_baseline
    strict = not args.relaxed_fp
    line = None
    if args.workload in ("large", "field"):
        line = measure_large(ctx, args.large_n * world, args.steps, args.warmup, field=args.workload == "field",
                             exchange=args.exchange, with_cpu=with_cpu, peaks=peaks)
    else:
        n = args.n_traj
        line = measure_lorenz(ctx, n, lorenz_ics(n, SEED + rank), args.steps, args.warmup, scaling="weak", peaks=peaks,
                              with_cpu=with_cpu, cpu_sample=args.cpu_sample, strict_fp=strict)
        if args.workload == "all":
            parity = None
            if world > 1:  # real multi-rank parity on this hardware, before the decomposed path is timed
                ok, how = check_slabs_against_single_slab(ctx, exchange=args.exchange)
                parity = {"slabs_equal_single_slab_bits_at_N_2^20": ok, "exchange": how}
                assert ok, "decomposed integration differs from the single-slab result"
            large = measure_large(ctx, args.large_n * world, args.steps, args.warmup, exchange=args.exchange,
                                  with_cpu=with_cpu, peaks=peaks)
            strong = None
            if world > 1:  # fixed total work split over the ranks, next to the weak numbers
                n_s = n // world
                ics = lorenz_ics(n, SEED)[rank * n_s:(rank + 1) * n_s]
                s_l = measure_lorenz(ctx, n_s, ics, args.steps, args.warmup, scaling="strong", peaks=peaks,
                                     with_cpu=False, cpu_sample=0)
                s_g = measure_large(ctx, args.large_n, args.steps, args.warmup, exchange=args.exchange, with_cpu=False,
                                    scaling="strong", peaks=peaks)
                strong = {"lorenz": s_l, "large": s_g}
            relaxed = None
            if world == 1:  # what the same kernel reaches with FMA contraction: labelled, opt-in, never the headline
                relaxed = measure_lorenz(ctx, n, lorenz_ics(n, SEED + rank), args.steps, args.warmup, scaling="weak",
                                         peaks=peaks, with_cpu=False, cpu_sample=0, strict_fp=False)
            if rank == 0:
                if parity:
                    large["parity_check"] = parity
                line["large"] = large
                if strong:
                    line["strong"] = strong
                if relaxed:
                    relaxed["note"] = ("opts.fp_mode = ODE_B200_FP_FMA: the same Dormand-Prince kernel compiled with FMA "
                                       "contraction; NOT bit-identical to the reference (within 1e-10 relative on "
                                       "non-chaotic problems, tests/test_gpu_ensemble.py), reported only to show the "
                                       "fraction of the FP64 FMA peak the design reaches when fusing is allowed")
                    line["relaxed_fp"] = relaxed
    if rank == 0:
        line["host"] = {"numa": numa, "ranks": world}
        print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
