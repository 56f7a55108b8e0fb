#!/usr/bin/env python3
"""bench.py -- headline benchmark: RK step attempts per second.

Workload (BASELINE.json configs[1]): Lorenz, ode45 (Dormand-Prince), 2^20
trajectories with randomised initial conditions, FP64, t in [0, 10], reference
default tolerances (reltol 1e-5, abstol 1e-8).  A "step" of this benchmark is
one pass of the hot path over the batch, i.e. the whole ensemble solve; the
metric counts RK step attempts (accepted + rejected, one attempt = one
rk_embedded_step! + stepsize_hw92!, /root/reference src/runge_kutta.jl:305-310).

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
Julia, which is not installed here or on the GPU box).  `--workload large`
benchmarks the single large reaction-diffusion system (configs[4]) instead.
Multi-GPU (torchrun): trajectories shard across ranks with no data-path
collective (weak scaling: 2^20 trajectories per GPU).
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
# dram__bytes_read.sum + dram__bytes_write.sum of one 2^20-trajectory launch, from the committed ncu capture
# (profiles/ensemble_lorenz_r1.md, capture prof_lorenz_r1_d: 60.71 MB + 14.20 MB; y0 25 MB + the pre-pass's
# {dt0, f0} 34 MB in, results out, most of which stay in L2); the kernel is register resident, traffic is I/O only
NCU_TRAFFIC_BYTES_PER_LAUNCH = 74.91e6
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


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port."""
    if rank != 0:
        return
    if args.workload in ("large", "field"):
        from bench_large import cpu_baseline_large
        cb = cpu_baseline_large()
        v = cb["attempts_per_s_at_2^26"]
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
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lorenz", choices=["lorenz", "large", "field"])
    ap.add_argument("--n-traj", type=int, default=1 << 20)
    ap.add_argument("--cpu-sample", type=int, default=1 << 18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload in ("large", "field"):
        from bench_large import run_large  # noqa: E402
        run_large(args, rank, local_rank, world)
        return

    import ctypes as C
    import torch
    import ode_jl_b200 as m
    from ode_jl_b200 import _lib

    L = m.lib()
    if not torch.cuda.is_available() or L.ode_b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = args.n_traj
    dof = 3
    y0_host = torch.from_numpy(lorenz_ics(n, SEED + rank)).pin_memory()
    F = m.DeviceRHS("lorenz", *LORENZ)
    tspan = np.array([0.0, 10.0])
    opts = m.make_opts("ode45", F, tspan, points="final", device=local_rank)

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    attempts = int((d_na.sum() + d_nr.sum()).item())
    assert int(d_st.max().item()) == 0, "a trajectory did not finish"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.ode_b200_launch_count(1)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:
        flush.fill_(1)  # L2 flush between timed iterations (not timed)
        e0.record(stream)
        step_dev()
        e1.record(stream)
    barrier()
    launches = int(L.ode_b200_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps

    # ---- end-to-end arm: host buffers through the C ABI ----------------------------
    out = dict(t_final=torch.empty(n, dtype=torch.float64).pin_memory(),
               y_final=torch.empty(n, dof, dtype=torch.float64).pin_memory(),
               n_accept=torch.empty(n, dtype=torch.int32).pin_memory(),
               n_reject=torch.empty(n, dtype=torch.int32).pin_memory(),
               status=torch.empty(n, dtype=torch.int32).pin_memory())
    par_h = np.array(LORENZ)
    io2 = _lib.EnsembleIO()
    io2.n_traj, io2.n_tspan, io2.trace_traj = n, 2, -1
    io2.y0, io2.params, io2.tspan = y0_host.data_ptr(), par_h.ctypes.data, tspan.ctypes.data
    io2.t_final, io2.y_final = out["t_final"].data_ptr(), out["y_final"].data_ptr()
    io2.n_accept, io2.n_reject, io2.status = out["n_accept"].data_ptr(), out["n_reject"].data_ptr(), out["status"].data_ptr()
    h2d = y0_host.numel() * 8 + 3 * 8 + 2 * 8
    d2h = n * 8 + n * dof * 8 + 3 * n * 4

    # One GPU per host: let the kernel write the results into the pinned output buffers itself (host_output = 1,
    # no copy-back phase: 12.4 -> 11.65 ms).  Several ranks on one host: staged copies -- with 8 B200s moving
    # 570 MB per step through one host the host side is the bottleneck either way (15.1 ms direct, 14.8 ms staged).
    opts_e2e = m.make_opts("ode45", F, tspan, points="final", device=local_rank, direct_output=(world == 1))

    def step_e2e():
        _lib.check(L.ode_b200_solve_ensemble(C.byref(opts_e2e), C.byref(io2)))

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    att2 = int(out["n_accept"].sum().item() + out["n_reject"].sum().item())
    assert att2 == attempts
    same = torch.equal(out["y_final"], d_yf.cpu())
    assert same, "host-buffer and device-resident arms disagree"

    # ---- reduce over ranks (max time, sum of work) ----------------------------------
    tot_attempts = attempts
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
        a = torch.tensor([attempts], dtype=torch.float64, device=dev)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        tot_attempts = int(a.item())

    if rank == 0:
        value = tot_attempts / (dev_ms * 1e-3)
        e2e_v = tot_attempts / e2e_s
        # roofline of the dominant kernel (this rank's launch): algorithmic flops / kernel time
        dfma = L.ode_b200_fp64_peak(local_rank, 0)   # FP64 FMA instr/s, measured now
        dmix = L.ode_b200_fp64_peak(local_rank, 1)   # DADD+DMUL instr/s, measured now
        achieved = attempts * FLOPS_PER_ATTEMPT / (dev_ms * 1e-3) / 1e12
        peak = 2.0 * dfma / 1e12
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "frac_of_issue_peak": achieved / (max(dfma, dmix) / 1e12),
                    "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH if n == (1 << 20) else None,
                    "peak_source": "measured live: ode_b200_fp64_peak DFMA chains x2 flops (no FP64 entry in "
                                   "MEASURED_PEAKS.json); issue peak = max(DFMA, DADD+DMUL) instr/s = %.3g" % max(dfma, dmix),
                    "note": "strict no-FMA arithmetic (reference parity) caps flops at 50% of the FMA peak; "
                            "frac_of_issue_peak counts algorithmic FP64 ops against FP64 instruction slots"}
        line = {"metric": "rk_step_attempts_per_s", "value": value, "unit": "steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "lorenz_ode45_dp_ensemble", "n_traj_per_gpu": n, "tspan": [0.0, 10.0],
                           "reltol": 1e-5, "abstol": 1e-8, "attempts_per_step_per_gpu": attempts,
                           "l2": "256 MiB flush write between timed iterations", "parallelism": "traj-shard x%d" % world},
                "e2e": {"output_path": "kernel writes pinned host buffers" if world == 1 else "staged D2H copies", "value": e2e_v, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_lorenz(args.cpu_sample)[0]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
