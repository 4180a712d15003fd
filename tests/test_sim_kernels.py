"""Pre-GPU debugging aid: the kernel SOURCES of libapmlcl compiled for the CPU with the emulation
shim in tests/sim/ (fibers for CUDA threads) and compared with the oracle.  This checks indexing,
boundary handling and the PCG/multigrid control flow without a GPU; it says nothing about the
product binary -- the GPU parity tests (tests/test_gpu_parity.py, -m gpu) do that through
libapmlcl.so.  The sim library is never loaded by the product package.
"""
import numpy as np
import pytest

from helpers import anchored_open_mask, load_map, synth_aperture
from oracle import oracle_binding as ob
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params, LEFT, RIGHT, TOP, BOTTOM
from sim import sim_lib


@pytest.fixture(scope="module")
def sim():
    return sim_lib.load()


def params(**kw):
    base = dict(inlet_pressure=100.0, outlet_pressure=0.0, density=1000.0, viscosity=1.0,
                voxel_size=20.0e-6, avg_factor=10.0, roughness=0.0, high_mask=100.0, low_mask=1e-4,
                outlet_side=TOP)
    base.update(kw)
    return Params.defaults(**base)


def check(sim, p, a, ptol=1e-6, qtol=1e-6, pmask=None):
    o, so = ob.solve(p, a)
    w, trx, trz, sc = lcl.conductance(p, a, lib=sim)
    assert np.array_equal(w, o["width"])
    assert np.array_equal(trx == 0, o["trx"] == 0) and np.array_equal(trz == 0, o["trz"] == 0)
    assert np.allclose(trx, o["trx"], rtol=1e-9, atol=0) and np.allclose(trz, o["trz"], rtol=1e-9, atol=0)
    assert (sc["zeros"], sc["masked"]) == (so["zeros"], so["masked"])
    f, s = lcl.lcl_solve(p, a, lib=sim)
    scale = np.abs(o["pressure"]).max()
    perr = np.abs(f["pressure"] - o["pressure"]) / scale
    if pmask is not None:       # see helpers.anchored_open_mask: floating pockets are ill defined in f64
        perr = perr[pmask]
    assert perr.max() < ptol
    assert s["flow"] == pytest.approx(so["flow"], rel=qtol)
    assert s["inlet_pressure"] == pytest.approx(so["inlet_pressure"], rel=1e-6, abs=1e-9)
    assert s["outlet_pressure"] == pytest.approx(so["outlet_pressure"], rel=1e-6, abs=1e-9)
    qscale = np.abs(o["flow_m"]).max()      # a component can be pure rounding noise (1-row maps)
    for k in ("flow_x", "flow_z", "flow_m"):
        assert np.abs(f[k] - o[k]).max() <= 1e-5 * qscale
    # REN = rho/mu QM / WIDTH: weighted by WIDTH this is the flux bar again (in a contact cell, WIDTH = the LOW-MASK
    # clamp, the division blows the flux noise of the solve up by 1e4 against the open cells that set max REN)
    assert np.abs((f["reynolds"] - o["reynolds"]) * o["width"]).max() <= 1e-5 * np.abs(o["reynolds"] * o["width"]).max()
    assert np.array_equal(f["aperture"], o["width"])
    assert s["converged"] == 1
    return f, s, o, so


@pytest.mark.parametrize("side", [LEFT, RIGHT, TOP, BOTTOM])
def test_sim_all_sides_odd_sizes(sim, side):
    a = load_map("Fracture1ApertureMap-10avg.txt")[20:73, 40:101]     # 53 x 61, has zero cells
    check(sim, params(outlet_side=side), a)


def test_sim_roughness_and_high_mask(sim):
    a = load_map("Fracture3ApertureMap-10avg.txt")[:48, :64]
    check(sim, params(roughness=0.25, high_mask=6.0), a, ptol=5e-6)


@pytest.mark.parametrize("mode", ["inlet_rate", "outlet_rate"])
def test_sim_rate_bc(sim, mode):
    a = load_map("Fracture3ApertureMap-10avg.txt")[30:70, 10:58]
    if mode == "inlet_rate":
        p = params(outlet_side=RIGHT, inlet_pressure=-1.0, outlet_pressure=50.0, inlet_rate=1e-9)
    else:
        p = params(outlet_side=BOTTOM, inlet_pressure=500.0, outlet_pressure=-1.0, outlet_rate=1e-9)
    f, s, o, so = check(sim, p, a)
    assert s["outlet_rate"] == pytest.approx(1e-9, rel=1e-6)


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (6, 1), (2, 2), (5, 9), (7, 7), (16, 3)])
def test_sim_tiny_and_degenerate_maps(sim, shape):
    a = synth_aperture(shape[0], shape[1], seed=shape[0] * 31 + shape[1], sigma_c=0)
    check(sim, params(voxel_size=1e-5, avg_factor=1.0, viscosity=1e-3, inlet_pressure=1000.0), a)


@pytest.mark.parametrize("shape", [(3, 1200), (900, 2), (1, 1500)])
def test_sim_thin_strip_maps(sim, shape):
    """1 x n tail levels: long framed rows in k_tail_cycle's shared memory."""
    a = synth_aperture(shape[0], shape[1], seed=shape[0] * 7 + shape[1], sigma_c=2)
    check(sim, params(voxel_size=1e-5, avg_factor=1.0, viscosity=1e-3, inlet_pressure=1000.0), a)


def test_sim_synthetic_contact_map(sim):
    a = synth_aperture(64, 80, seed=5, contact=0.2, sigma_k=5)
    f, s, o, so = check(sim, params(voxel_size=1e-5, avg_factor=1.0, viscosity=1e-3, inlet_pressure=1000.0), a,
                        pmask=anchored_open_mask(a))
    assert so["zeros"] > 0


@pytest.mark.parametrize("base", [1e7, 1e9])      # 1e5 as well in the GPU variant (tests/test_gpu_parity.py)
def test_sim_back_pressure_does_not_cost_flux_accuracy(sim, base):
    """A common pressure offset must not change the fluxes (they are pressure differences): the solve runs on
    the unit drop and is shifted afterwards.  Solving for absolute pressures lost 1e-3 of the local fluxes at
    1e7 + 100 -> 1e7 Pa while reporting convergence (round-1 advisor finding)."""
    a = synth_aperture(96, 80, seed=3, sigma_c=4, contact=0.1, sigma_k=6)
    kw = dict(voxel_size=1e-5, avg_factor=1.0, viscosity=1e-3)
    f0, s0 = lcl.lcl_solve(params(inlet_pressure=100.0, outlet_pressure=0.0, **kw), a, lib=sim)
    f1, s1 = lcl.lcl_solve(params(inlet_pressure=base + 100.0, outlet_pressure=base, **kw), a, lib=sim)
    assert s1["converged"] == 1 and s1["iterations"] == s0["iterations"]
    assert s1["flow"] == pytest.approx(s0["flow"], rel=1e-6)
    # what is left is the rounding of the stored absolute pressures (eps * base per cell), which the
    # reference's FRAC_FLOW sees as well: the bar is 1e-6 of the largest flux plus that floor
    floor = 4 * np.finfo(float).eps * base / (100.0 / 96)     # ulp of the stored pressures vs the pressure step per cell row
    for k in ("flow_x", "flow_z"):
        assert np.abs(f1[k] - f0[k]).max() <= (1e-6 + floor) * np.abs(f0["flow_m"]).max()
    assert np.abs((f1["pressure"] - base) - f0["pressure"]).max() <= (1e-6 + floor) * 100.0


def test_sim_bulk_sweep_shares_unit_solutions(sim, monkeypatch):
    from helpers import check_bulk_sweep_shares_unit_solutions
    kw = dict(voxel_size=1e-5, avg_factor=1.0, viscosity=1e-3)
    check_bulk_sweep_shares_unit_solutions(lcl, sim, lambda **k: params(**dict(kw, **k)), monkeypatch, shape=(20, 18), full=False)


def test_sim_invalid_boundary_conditions_fail(sim):
    a = synth_aperture(8, 8, seed=1, sigma_c=0)
    with pytest.raises(lcl._lib.LclError):
        lcl.lcl_solve(params(inlet_rate=1e-9), a, lib=sim)          # both pressures and a rate
    with pytest.raises(lcl._lib.LclError):
        lcl.lcl_solve(params(viscosity=-1.0), a, lib=sim)


def _solver_vector(solver, name):
    import ctypes as C
    ptr, cnt = C.c_void_p(), C.c_size_t()
    assert solver.lib.apm_lcl_solver_device_field(solver._h, name.encode(), C.byref(ptr), C.byref(cnt)) == 0
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(cnt.value,))   # sim: host memory


@pytest.mark.parametrize("w_levels", [None, (1, 2)])
def test_sim_preconditioner_is_symmetric_positive(sim, monkeypatch, w_levels):
    """CG needs M = M^T > 0.  The cycle is assembled from fused, tile-local kernels (down leg, up leg,
    hybrid 9-point smoothers, optionally the truncated W-cycle): apply it to two vectors through the
    solver's scratch arrays and compare <M a, b> with <a, M b>.  170 x 150 spans several tiles of every
    kernel in both directions, with ragged edges."""
    if w_levels:
        monkeypatch.setenv("APM_LCL_W_LO", str(w_levels[0]))
        monkeypatch.setenv("APM_LCL_W_HI", str(w_levels[1]))
    nz, nx = 170, 150
    a = synth_aperture(nz, nx, seed=5, sigma_c=4, contact=0.2, sigma_k=6)
    p = params(viscosity=1e-3, voxel_size=1e-5, avg_factor=1.0)
    with lcl.LclSolver(nx, nz, lib=sim) as solver:
        solver.solve(p, a, fields=("pressure",))
        tx, tz = _solver_vector(solver, "tx").reshape(nz, nx), _solver_vector(solver, "tz").reshape(nz, nx)
        txl, tzb = _solver_vector(solver, "txl"), _solver_vector(solver, "tzb")
        diag = (np.concatenate([txl[:, None], tx[:, :-1]], axis=1) + tx) + (np.concatenate([tzb[None, :], tz[:-1]], axis=0) + tz)
        r, q = _solver_vector(solver, "r"), _solver_vector(solver, "q")
        rng = np.random.default_rng(1)
        va, vb = (diag * rng.standard_normal((nz, nx))).ravel(), (diag * rng.standard_normal((nz, nx))).ravel()

        def apply(v):
            r[:] = v
            solver.time_kernel("vcycle", reps=1)
            return q.copy()
        ma, mb = apply(va), apply(vb)
    aa, bb, ab, ba = ma @ va, mb @ vb, ma @ vb, mb @ va
    assert aa > 0 and bb > 0
    assert abs(ab - ba) <= 1e-10 * np.sqrt(aa * bb)


def test_sim_bulk_solve_reports_failed_cases_and_keeps_the_others(sim):
    """apm_lcl_bulk_solve returns the number of failed cases, marks them converged = 0, keeps the first message and still
    delivers every other case (include/apm_lcl.h); a case that runs out of iterations inside a group of shared unit
    solutions must not poison the cases after it."""
    import ctypes as C
    from apmapflow_b200._lib import Fields, Scalars
    a = synth_aperture(24, 20, seed=5, sigma_c=3, contact=0.1, sigma_k=5)
    kw = dict(voxel_size=1e-5, avg_factor=1.0, viscosity=1e-3)
    cases = [params(**kw), params(**dict(kw, viscosity=-1.0)), params(inlet_rate=1e-9, **kw),          # ok, bad, bad
             params(outlet_pressure=250.0, **kw), params(max_iter=1, **kw), params(outlet_pressure=500.0, **kw)]
    n = len(cases)
    arr = (Params * n)(*cases)
    maps = (C.c_void_p * n)(*[a.ctypes.data] * n)
    outs = (Fields * n)()
    pres = [np.full(a.shape, np.nan) for _ in range(n)]
    for k in range(n):
        outs[k].pressure = pres[k].ctypes.data
    scal = (Scalars * n)()
    failed = sim.apm_lcl_bulk_solve(0, a.shape[1], a.shape[0], n, arr, maps, outs, scal)
    assert failed == 3
    assert [scal[k].converged for k in range(n)] == [1, 0, 0, 1, 0, 1]
    assert b"VISCOSITY" in sim.apm_lcl_last_error()          # the message of the first failed case
    for k, p_out in ((0, 0.0), (3, 250.0), (5, 500.0)):
        f, s = lcl.lcl_solve(cases[k], a, fields=("pressure",), lib=sim)
        assert np.abs(pres[k] - f["pressure"]).max() <= 1e-9 * 100.0
        assert scal[k].flow == pytest.approx(s["flow"], rel=1e-8) and scal[k].outlet_pressure == p_out
    assert scal[3].iterations == 0 and scal[5].iterations == 0      # derived from case 0's unit solution
