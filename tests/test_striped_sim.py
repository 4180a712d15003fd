"""One map striped over two ranks on the CPU: two gloo processes, each running the kernel SOURCES
of libapmlcl through the emulation shim in tests/sim/ on a POSIX shared-memory slab in place of the CUDA IPC
slab.  Checks the whole distributed control flow -- strips, ghost-row stores and flags, gathers of the dot
products, the broadcast onto the first whole level -- against the single-rank solve of the same map, which
the other sim tests pin to the oracle.  The product path (libapmlcl.so, CUDA IPC over NVLink) is tested by
tests/test_gpu_striped.py on the GPU box."""
import os
import pickle

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import synth_aperture
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params, LEFT, RIGHT, TOP, BOTTOM


def params(**kw):
    base = dict(inlet_pressure=1000.0, outlet_pressure=0.0, density=1000.0, viscosity=1e-3, voxel_size=1e-5,
                avg_factor=1.0, roughness=0.0, high_mask=100.0, low_mask=1e-4, outlet_side=TOP)
    base.update(kw)
    return base


def _worker(rank, world, port, tmpdir, shape, seed, pkw, env):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), **env)
    from sim import sim_lib
    sim = sim_lib.load()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nz, nx = shape
        a = synth_aperture(nz, nx, seed=seed, sigma_c=4, contact=0.15, sigma_k=6)
        p = Params.defaults(**pkw)
        with lcl.StripedLclSolver(nx, nz, lib=sim) as solver:
            f, s = solver.solve(p, a)
            s["rows"] = (solver.z_lo, solver.z_hi)
            import ctypes as C
            cnt = C.c_size_t()
            ptr = C.c_void_p()
            sim.apm_lcl_solver_device_field(solver._h, b"gather_level", C.byref(ptr), C.byref(cnt))
            s["gather_level"] = cnt.value
            whole = solver.gather(f)
        allsc = [None] * world
        dist.all_gather_object(allsc, s)
        if rank == 0:
            with open(os.path.join(tmpdir, "out.pkl"), "wb") as fh:
                pickle.dump((whole, allsc), fh)
    finally:
        dist.destroy_process_group()


def run_striped(tmp_path, world, shape, seed, pkw, min_cells):
    port = 29500 + (os.getpid() * 7 + shape[0]) % 2000
    env = {"APM_LCL_STRIPE_MIN_CELLS": str(min_cells)}
    mp.spawn(_worker, args=(world, port, str(tmp_path), shape, seed, pkw, env), nprocs=world, join=True)
    with open(tmp_path / "out.pkl", "rb") as fh:
        return pickle.load(fh)


def single(shape, seed, pkw):
    from sim import sim_lib
    nz, nx = shape
    a = synth_aperture(nz, nx, seed=seed, sigma_c=4, contact=0.15, sigma_k=6)
    return lcl.lcl_solve(Params.defaults(**pkw), a, lib=sim_lib.load())


def compare(whole, allsc, f1, s1, world):
    # every rank holds the same whole-map scalars, bit for bit
    for k in ("flow", "zeros", "masked", "raw_aper", "avg_aper", "min_aper", "max_aper", "iterations",
              "rel_residual", "max_residual", "inlet_pressure", "outlet_pressure"):
        assert len({sc[k] for sc in allsc}) == 1, k
    s = allsc[0]
    assert s["converged"] == 1
    assert (s["zeros"], s["masked"], s["min_aper"], s["max_aper"]) == (s1["zeros"], s1["masked"], s1["min_aper"], s1["max_aper"])
    # the strips run the single-rank iteration (same smoother tiles, exact fine sweeps); only the order of the
    # dot-product sums differs
    assert abs(s["iterations"] - s1["iterations"]) <= 1
    assert s["flow"] == pytest.approx(s1["flow"], rel=1e-9)
    scale = np.abs(f1["pressure"]).max()
    assert np.abs(whole["pressure"] - f1["pressure"]).max() <= 2e-7 * scale
    assert np.array_equal(whole["aperture"], f1["aperture"])
    for k in ("flow_x", "flow_z", "flow_m"):
        assert np.abs(whole[k] - f1[k]).max() <= 1e-6 * np.abs(f1["flow_m"]).max()
    assert np.nanmax(whole["residual"]) == 1.0


@pytest.mark.parametrize("side", [TOP, LEFT])
def test_two_strips_three_striped_levels(tmp_path, side):
    """256 x 80 map, two ranks, levels 0-2 striped (cut at a multiple of 64 rows), level 3 -- where the
    single-block tail cycle starts -- held whole by both."""
    shape, pkw = (256, 80), params(outlet_side=side)
    whole, allsc = run_striped(tmp_path, 2, shape, 3, pkw, min_cells=500)
    assert allsc[0]["gather_level"] == 3 and [sc["rows"] for sc in allsc] == [(0, 128), (128, 256)]
    f1, s1 = single(shape, 3, pkw)
    compare(whole, allsc, f1, s1, 2)


def test_w_cycle_on_striped_levels(tmp_path, monkeypatch):
    """Levels 1-2 visited twice per cycle (the truncated W-cycle of large maps): the second visit starts from
    the first one's result, whose ghost rows must have crossed the cut."""
    monkeypatch.setenv("APM_LCL_W_LO", "1")
    monkeypatch.setenv("APM_LCL_W_HI", "2")
    shape, pkw = (128, 96), params()
    whole, allsc = run_striped(tmp_path, 2, shape, 4, pkw, min_cells=500)
    assert allsc[0]["gather_level"] == 2
    f1, s1 = single(shape, 4, pkw)
    compare(whole, allsc, f1, s1, 2)


def test_three_ragged_strips_rate_bc(tmp_path):
    """Three ranks, a map whose height is no multiple of anything, a flow-rate boundary condition (side flow
    summed over the ranks) on the BOTTOM outlet."""
    shape = (150, 70)
    pkw = params(outlet_side=BOTTOM, inlet_pressure=500.0, outlet_pressure=-1.0, outlet_rate=1e-9)
    whole, allsc = run_striped(tmp_path, 3, shape, 5, pkw, min_cells=2000)
    assert allsc[0]["gather_level"] == 2
    assert [sc["rows"] for sc in allsc] == [(0, 64), (64, 96), (96, 150)]   # cuts at multiples of 32 rows
    f1, s1 = single(shape, 5, pkw)
    compare(whole, allsc, f1, s1, 3)
    assert allsc[0]["outlet_rate"] == pytest.approx(1e-9, rel=1e-6)


def test_odd_width_and_inlet_rate_on_the_left(tmp_path):
    """Odd nx (no paired 128-bit accesses, no interior-tile fast path) and a group-node inlet on a LEFT/RIGHT pair:
    the side flow is then a sum over every rank's rows."""
    shape = (128, 75)
    pkw = params(outlet_side=RIGHT, inlet_pressure=-1.0, outlet_pressure=50.0, inlet_rate=1e-9)
    whole, allsc = run_striped(tmp_path, 2, shape, 9, pkw, min_cells=500)
    assert allsc[0]["gather_level"] == 2
    f1, s1 = single(shape, 9, pkw)
    compare(whole, allsc, f1, s1, 2)
    assert allsc[0]["outlet_rate"] == pytest.approx(1e-9, rel=1e-6)
    assert allsc[0]["inlet_pressure"] == pytest.approx(s1["inlet_pressure"], rel=1e-9)


def test_map_too_small_to_stripe_is_solved_whole_by_every_rank(tmp_path):
    shape, pkw = (40, 48), params()
    whole, allsc = run_striped(tmp_path, 2, shape, 1, pkw, min_cells=1 << 20)
    assert allsc[0]["gather_level"] == 0 and [sc["rows"] for sc in allsc] == [(0, 40), (0, 40)]
    f1, s1 = single(shape, 1, pkw)
    assert np.array_equal(whole["pressure"], f1["pressure"]) and allsc[1]["flow"] == s1["flow"]
