"""The N>1 path on CPU: world_size-2 (and 3) gloo process groups drive the same host loop
(`integrate_slab`) that the CUDA ranks run, with a CPU stand-in slab engine.  Checks the
sharding helper, the all-gather layout of the exchange buffer, identical decisions on every
rank, and that the decomposed result equals the oracle's single-process result bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _u0(n):
    i = np.arange(n)
    return 0.5 + 0.4 * np.sin(2 * np.pi * 8 * i / n)


def _worker(rank, world, port, n, method, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ode_jl_b200 as m
    from cpu_slab_engine import CpuSlabEngine
    lo, hi = m.shard_bounds(n, world, rank)
    send = torch.zeros(m._lib.XCHG_DOUBLES, dtype=torch.float64)
    recv = torch.zeros(world, m._lib.XCHG_DOUBLES, dtype=torch.float64)
    eng = CpuSlabEngine(method, (1.0, 1.0), hi - lo, rank, world, [0.0, 2.0], send.numpy(), recv.numpy())
    eng.set_state(_u0(n)[lo:hi])

    def exchange():
        dist.all_gather_into_tensor(recv.view(-1), send)

    ctl = m.integrate_slab(eng, 0, 0, exchange, poll_every=3)
    q.put((rank, lo, hi, eng.get_state(), ctl["n_accept"], ctl["n_reject"], ctl["t_final"], eng.trace))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n, method):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + world
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, method, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=180) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world,method", [(2, "ode45"), (3, "ode23")])
def test_gloo_slab_driver_matches_oracle(world, method):
    from oracle import oracle as O
    n = 331 if world == 3 else 256
    res = _run(world, n, method)
    y = np.concatenate([r[3] for r in res])
    assert [(r[1], r[2]) for r in res][0][0] == 0 and res[-1][2] == n
    # every rank took the same decisions
    assert len({(r[4], r[5], r[6]) for r in res}) == 1
    assert all(r[7] == res[0][7] for r in res)
    ref = O.solve(method, "reaction_diffusion", _u0(n), [0.0, 2.0], [1.0, 1.0], trace=True, points="final")
    assert (res[0][4], res[0][5]) == (ref.n_accept, ref.n_reject)
    assert np.array_equal(np.array(res[0][7]), ref.trace)
    assert np.array_equal(y, ref.y[-1])


def test_shard_bounds_cover_the_grid():
    import ode_jl_b200 as m
    for n, w in ((10, 3), (1 << 20, 8), (17, 17), (1000, 7)):
        b = [m.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _ens_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ode_jl_b200 as m
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
    mu = rng.uniform(0.5, 2.0, (n, 3)) * np.array([10.0, 28.0, 8 / 3])

    def cpu_solve(alg, F, y, tspan, params=None, **kw):  # stand-in for the CUDA shard solve (tests only)
        return O.solve_ensemble(alg, F.name, y, tspan, params, threads=1, **kw)

    F = m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3)
    full = m.solve_ensemble_distributed("ode45", F, y0, [0.0, 0.5], params=mu, _local_solve=cpu_solve)
    part = m.solve_ensemble_distributed("ode45", F, y0, [0.0, 0.5], params=mu, gather=False, _local_solve=cpu_solve)
    q.put((rank, full, part))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ensemble_shards_and_final_gather(world):
    """Trajectory shards + the final gather (the only communication of the ensemble regime)."""
    from oracle import oracle as O
    n = 101
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200) + world
    ps = [ctx.Process(target=_ens_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=180) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
    mu = rng.uniform(0.5, 2.0, (n, 3)) * np.array([10.0, 28.0, 8 / 3])
    ref = O.solve_ensemble("ode45", "lorenz", y0, [0.0, 0.5], mu)
    for rank, full, part in res:
        for k in ("t_final", "y_final", "n_accept", "n_reject", "n_rhs", "status"):
            assert np.array_equal(full[k], ref[k]), (rank, k)
            lo, hi = part["rows"]
            assert np.array_equal(part[k], ref[k][lo:hi])
    assert [r[2]["rows"] for r in res][0][0] == 0 and res[-1][2]["rows"][1] == n
