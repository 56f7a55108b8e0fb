"""Multi-GPU fan-out INSIDE the library (one process, opts.n_gpus > 1 / ode_b200_solve_large_multi): needs >= 2
devices, skipped otherwise (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

Ensembles: contiguous trajectory blocks, one host thread per device, no collective.  Large system: slabs whose
attempt kernels exchange their records over NVLink peer memory and run the step controller themselves; the result
must be the single-GPU result (and the oracle's) bit for bit."""
import ctypes as C

import numpy as np
import pytest

import ode_jl_b200 as m
from ode_jl_b200 import _lib
from oracle import oracle as O

pytestmark = pytest.mark.gpu
RD = [1.0, 1.0]


def n_dev():
    return m.lib().ode_b200_device_count()


needs2 = pytest.mark.skipif(n_dev() < 2, reason="needs >= 2 CUDA devices")


def u0(n):
    return 0.5 + 0.4 * np.sin(2 * np.pi * 8 * np.arange(n) / n)


def gpus_to_try():
    return [g for g in (2, 3, 4, 8) if g <= n_dev()]


@needs2
@pytest.mark.parametrize("n", [30011, 1 << 20])
def test_large_system_on_n_gpus_equals_one_gpu_and_oracle(n):
    y0 = u0(n)
    F = m.DeviceRHS("reaction_diffusion", *RD)
    one = m.solve_large("ode45", F, y0, [0.0, 20.0], trace=True)
    if n < 100000:
        ref = O.solve("ode45", "reaction_diffusion", y0, [0.0, 20.0], RD, trace=True, points="final")
        assert np.array_equal(one["y"], ref.y[-1]) and np.array_equal(one["trace"], ref.trace)
    for g in gpus_to_try():
        got = m.solve_large("ode45", F, y0, [0.0, 20.0], trace=True, n_gpus=g)
        assert got["status"] == 0, (g, got["status"])
        assert (got["n_accept"], got["n_reject"], got["n_rhs"]) == (one["n_accept"], one["n_reject"], one["n_rhs"])
        assert np.array_equal(got["trace"], one["trace"]), g
        assert np.array_equal(got["y"], one["y"]), g
        again = m.solve_large("ode45", F, y0, [0.0, 20.0], n_gpus=g)  # the cached sessions re-arm cleanly
        assert np.array_equal(again["y"], one["y"]), g


@needs2
@pytest.mark.parametrize("method", ["ode23", "ode45_fe", "ode78"])
def test_other_tableaus_on_two_gpus(method):
    n = 20011
    y0 = u0(n) + 0.05 * np.cos(2 * np.pi * 3 * np.arange(n) / n)
    F = m.DeviceRHS("reaction_diffusion", 0.7, 1.3)
    one = m.solve_large(method, F, y0, [0.0, 3.0], trace=True)
    two = m.solve_large(method, F, y0, [0.0, 3.0], trace=True, n_gpus=2)
    assert np.array_equal(two["trace"], one["trace"]) and np.array_equal(two["y"], one["y"])
    assert (two["n_accept"], two["n_reject"], two["n_rhs"]) == (one["n_accept"], one["n_reject"], one["n_rhs"])


@needs2
def test_literal_repeat_of_a_nonfinite_attempt_on_two_gpus():
    """initstep far too long: the stages overflow, every device asks for the literal repeat from the same gathered
    flags, the NaN guard cuts dt and the run recovers -- same trace as one GPU and as the oracle."""
    n = 6000
    y0 = u0(n)
    F = m.DeviceRHS("reaction_diffusion", *RD)
    kw = dict(initstep=1e7)
    ref = O.solve("ode45", "reaction_diffusion", y0, [0.0, 1.0], RD, trace=True, points="final", **kw)
    two = m.solve_large("ode45", F, y0, [0.0, 1.0], trace=True, n_gpus=2, **kw)
    assert ref.status == 0 and np.any(ref.trace[:, 2] == 10.0)
    assert two["status"] == 0 and np.array_equal(two["trace"], ref.trace, equal_nan=True)
    assert np.array_equal(two["y"], ref.y[-1])


@needs2
def test_ensemble_fan_out_equals_oracle_and_one_gpu():
    n = 20001
    rng = np.random.default_rng(5)
    y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
    F = m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3)
    one = m.solve_ensemble("ode45", F, y0, [0.0, 2.0])
    idx = rng.choice(n, 512, replace=False)
    ref = O.solve_ensemble("ode45", "lorenz", y0[idx], [0.0, 2.0], [10.0, 28.0, 8 / 3])
    for g in gpus_to_try():
        got = m.solve_ensemble("ode45", F, y0, [0.0, 2.0], n_gpus=g)
        for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
            assert np.array_equal(got[k], one[k]), (g, k)
            assert np.array_equal(got[k][idx], ref[k]), (g, k)
    # per-trajectory parameters and points output are sliced with the trajectories
    mu = rng.uniform(0.5, 3.0, (999, 1))
    v0 = rng.uniform(-2, 2, (999, 2))
    V = m.DeviceRHS("vdp", 1.0)
    ts = [0.0, 0.5, 1.0, 2.0]
    a = m.solve_ensemble("ode23", V, v0, ts, params=mu, points="specified")
    b = m.solve_ensemble("ode23", V, v0, ts, params=mu, points="specified", n_gpus=2)
    for k in ("y_final", "pts_y", "pts_t", "pts_n", "n_accept"):
        assert np.array_equal(a[k], b[k]), k


@needs2
def test_multi_error_paths():
    F = m.DeviceRHS("reaction_diffusion", *RD)
    with pytest.raises(m.OdeB200Error):
        m.solve_large("ode45", F, u0(4096), [0.0, 1.0], n_gpus=n_dev() + 1)
    with pytest.raises(m.OdeB200Error):  # slabs shorter than two ghost zones (ode78: H = 12)
        m.solve_large("ode78", F, u0(40), [0.0, 1.0], n_gpus=2)
    ok = m.solve_large("ode45", F, u0(40), [0.0, 1.0], n_gpus=2)  # 20 points per slab, H = 8: fine
    assert np.array_equal(ok["y"], m.solve_large("ode45", F, u0(40), [0.0, 1.0])["y"])
