"""GPU parity tests proper (-m gpu): libapmlcl.so through its C ABI vs the CPU oracle, the
reference's golden files, and size-independent properties at BASELINE.json's full sizes.

Tolerances (north_star): masks / counters / cell indexing bit exact; pressure within 1e-6 of
max|p|; flow rate within 1e-6 relative.  Face conductances are compared at 1e-9 relative because
CUDA's atan/tan differ from glibc's in the last ulp and tan(theta)-theta cancels.
"""
import numpy as np
import pytest

from helpers import anchored_open_mask, load_csv, load_map, synth_aperture
from oracle import oracle_binding as ob
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params, LEFT, RIGHT, TOP, BOTTOM

pytestmark = pytest.mark.gpu


def params(**kw):
    base = dict(inlet_pressure=100.0, outlet_pressure=0.0, density=1000.0, viscosity=1.0,
                voxel_size=20.0e-6, avg_factor=10.0, roughness=0.0, high_mask=100.0, low_mask=1e-4,
                outlet_side=TOP)
    base.update(kw)
    return Params.defaults(**base)


def synth_params(**kw):
    # SURVEY.md 8d common .inp values
    base = dict(inlet_pressure=1000.0, outlet_pressure=0.0, density=1000.0, viscosity=1.0e-3,
                voxel_size=10.0e-6, avg_factor=1.0, roughness=0.0, high_mask=100.0, low_mask=1e-4,
                outlet_side=TOP)
    base.update(kw)
    return Params.defaults(**base)


def check_against_oracle(p, a, ptol=1e-6, qtol=1e-6, pmask=None):
    o, so = ob.solve(p, a)
    w, trx, trz, sc = lcl.conductance(p, a)
    assert np.array_equal(w, o["width"])                                  # clamp masks: bit exact
    assert np.array_equal(trx == 0, o["trx"] == 0) and np.array_equal(trz == 0, o["trz"] == 0)
    assert np.allclose(trx, o["trx"], rtol=1e-9, atol=0) and np.allclose(trz, o["trz"], rtol=1e-9, atol=0)
    assert (sc["zeros"], sc["masked"]) == (so["zeros"], so["masked"])     # counters: bit exact
    assert sc["min_aper"] == so["min_aper"] and sc["max_aper"] == so["max_aper"]
    assert sc["raw_aper"] == pytest.approx(so["raw_aper"], rel=1e-12)
    assert sc["avg_aper"] == pytest.approx(so["avg_aper"], rel=1e-12)
    f, s = lcl.lcl_solve(p, a)
    assert s["gpu_launches"] > 0 and s["converged"] == 1
    scale = np.abs(o["pressure"]).max()
    perr = np.abs(f["pressure"] - o["pressure"]) / scale
    if pmask is not None:       # see helpers.anchored_open_mask: floating pockets are ill defined in f64
        perr = perr[pmask]
    assert perr.max() < ptol
    assert s["flow"] == pytest.approx(so["flow"], rel=qtol)
    assert s["inlet_pressure"] == pytest.approx(so["inlet_pressure"], rel=1e-6, abs=1e-9)
    assert s["outlet_pressure"] == pytest.approx(so["outlet_pressure"], rel=1e-6, abs=1e-9)
    qscale = np.abs(o["flow_m"]).max()
    for k in ("flow_x", "flow_z", "flow_m"):
        assert np.abs(f[k] - o[k]).max() <= 1e-5 * qscale
    # REN = rho/mu QM / WIDTH: weighted by WIDTH this is the flux bar again (in a contact cell, WIDTH = the LOW-MASK
    # clamp, the division blows the flux noise of the solve up by 1e4 against the open cells that set max REN)
    assert np.abs((f["reynolds"] - o["reynolds"]) * o["width"]).max() <= 1e-5 * np.abs(o["reynolds"] * o["width"]).max()
    assert np.array_equal(f["aperture"], o["width"])
    r = f["residual"]
    if not np.isnan(r).all():          # every Q5 == 0 gives 0/0 = NaN, exactly like the reference (:826)
        assert np.nanmax(r) == 1.0 and np.nanmin(r) >= 0.0
    return f, s, o, so


def test_golden_parallel_plate_matches_reference_files():
    """BASELINE config 1: coverage_test.inp on ParallelPlate-1vox.txt vs the reference's own outputs."""
    a = load_map("ParallelPlate-1vox.txt")
    p = params(voxel_size=10e-6)
    f, s = lcl.lcl_solve(p, a)
    assert np.abs(f["pressure"] - load_csv("coverage_test_pres-orig.csv.gz")).max() <= 1e-6 * 100.0
    assert np.allclose(f["flow_z"], load_csv("coverage_test_flow-z-orig.csv.gz"), rtol=1e-6, atol=0)
    assert np.allclose(f["flow_m"], load_csv("coverage_test_flow-m-orig.csv.gz"), rtol=1e-6, atol=0)
    assert np.allclose(f["aperture"], load_csv("coverage_test_aper-orig.csv.gz"), rtol=1e-7, atol=0)
    assert np.abs(f["flow_x"]).max() < 1e-6 * np.abs(f["flow_z"]).max()
    assert s["outlet_rate"] == pytest.approx(0.833333333333e-14, rel=1e-6)      # coverage_test_stats-orig.yaml
    assert (s["zeros"], s["masked"]) == (0, 0)


@pytest.mark.parametrize("name", ["Fracture1ApertureMap-10avg.txt", "Fracture3ApertureMap-10avg.txt"])
@pytest.mark.parametrize("side", [LEFT, RIGHT, TOP, BOTTOM])
def test_fixture_maps_all_outlet_sides(name, side):
    check_against_oracle(params(outlet_side=side), load_map(name))


def test_roughness_and_high_mask():
    check_against_oracle(params(roughness=0.25, high_mask=6.0), load_map("Fracture3ApertureMap-10avg.txt"), ptol=5e-6)


@pytest.mark.parametrize("side", [LEFT, TOP])
@pytest.mark.parametrize("mode", ["inlet_rate", "outlet_rate"])
def test_rate_boundary_conditions(side, mode):
    a = load_map("Fracture3ApertureMap-10avg.txt")
    if mode == "inlet_rate":
        p = params(outlet_side=side, inlet_pressure=-1.0, outlet_pressure=50.0, inlet_rate=1e-9)
    else:
        p = params(outlet_side=side, inlet_pressure=500.0, outlet_pressure=-1.0, outlet_rate=1e-9)
    f, s, o, so = check_against_oracle(p, a)
    assert s["outlet_rate"] == pytest.approx(1e-9, rel=1e-6)


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (6, 1), (2, 2), (5, 9), (7, 7), (16, 3), (33, 65), (100, 37)])
def test_tiny_ragged_and_degenerate_maps(shape):
    a = synth_aperture(shape[0], shape[1], seed=shape[0] * 31 + shape[1], sigma_c=0)
    check_against_oracle(synth_params(), a)


@pytest.mark.parametrize("n,contact", [(256, 0.0), (256, 0.2), (384, 0.2)])
def test_synthetic_lognormal_maps_vs_oracle(n, contact):
    """BASELINE configs 2/3 recipe at sizes the direct D4 oracle finishes in seconds."""
    a = synth_aperture(n, n, seed=n, sigma_c=8, contact=contact, sigma_k=16)
    f, s, o, so = check_against_oracle(synth_params(), a, pmask=anchored_open_mask(a) if contact else None)
    if contact:
        assert so["zeros"] > 0.15 * n * n


@pytest.mark.parametrize("shape", [(131, 257), (173, 119), (59, 175), (86, 61)])
@pytest.mark.parametrize("side", [LEFT, RIGHT, TOP, BOTTOM])
def test_ragged_multi_tile_maps_all_sides_vs_oracle(shape, side):
    """Shapes that are no multiple of any kernel's tile (42x58 / 28x60 fine legs, 16x64 coarse tiles), odd
    and even widths (the paired double2 stores need an even nx), every outlet side, with contact cells."""
    nz, nx = shape
    a = synth_aperture(nz, nx, seed=nz + nx, sigma_c=4, contact=0.15, sigma_k=6)
    vertical = side in (TOP, BOTTOM)
    check_against_oracle(synth_params(outlet_side=side), a, pmask=anchored_open_mask(a, 1e-4, vertical))


@pytest.mark.parametrize("shape", [(3, 3000), (2500, 2), (5, 1500), (1, 4000)])
@pytest.mark.parametrize("side", [TOP, LEFT])
def test_thin_strip_maps_vs_oracle(shape, side):
    """Strips a few cells wide: the hierarchy degenerates to 1 x n levels, the levels k_tail_cycle holds in
    shared memory are long framed rows (the shapes its shared-memory budget is sized for)."""
    nz, nx = shape
    a = synth_aperture(nz, nx, seed=nz * 7 + nx, sigma_c=2)
    check_against_oracle(synth_params(outlet_side=side), a)


@pytest.mark.parametrize("base", [1e5, 1e7, 1e9])
def test_back_pressure_does_not_cost_flux_accuracy(base):
    """Fluxes are pressure differences: a common offset of the two boundary pressures must leave them alone
    (the solve runs on the unit drop and is shifted afterwards; round-1 advisor finding)."""
    a = synth_aperture(300, 260, seed=3, sigma_c=6, contact=0.1, sigma_k=10)
    f0, s0 = lcl.lcl_solve(synth_params(inlet_pressure=100.0, outlet_pressure=0.0), a)
    f1, s1 = lcl.lcl_solve(synth_params(inlet_pressure=base + 100.0, outlet_pressure=base), a)
    assert s1["converged"] == 1 and s1["iterations"] == s0["iterations"]
    assert s1["flow"] == pytest.approx(s0["flow"], rel=1e-6)
    floor = 4 * np.finfo(float).eps * base / (100.0 / 300)    # rounding of the stored absolute pressures vs the pressure step per cell row
    for k in ("flow_x", "flow_z"):
        assert np.abs(f1[k] - f0[k]).max() <= (1e-6 + floor) * np.abs(f0["flow_m"]).max()
    assert np.abs((f1["pressure"] - base) - f0["pressure"]).max() <= (1e-6 + floor) * 100.0


def test_invalid_inputs_fail_with_reference_messages():
    a = synth_aperture(8, 8, seed=1, sigma_c=0)
    with pytest.raises(lcl._lib.LclError, match="INVALID BOUNDARY"):
        lcl.lcl_solve(params(inlet_rate=1e-9), a)
    with pytest.raises(lcl._lib.LclError, match="INVAILD VISCOSITY"):
        lcl.lcl_solve(params(viscosity=-1.0), a)


def _inflow_outflow(f, p_in, p_out, trz):
    q_in = (trz[0, :] * (p_in - f["pressure"][0, :])).sum()
    q_out = (trz[-1, :] * (f["pressure"][-1, :] - p_out)).sum()
    return q_in, q_out


@pytest.mark.parametrize("n,contact", [(2048, 0.0), (2048, 0.2)])
def test_full_size_properties_2048(n, contact):
    """BASELINE config 2 at full size: no oracle possible in seconds, so size-independent checks:
    mass balance, maximum principle, self-convergence under a 1000x tighter tolerance, determinism."""
    a = synth_aperture(n, n, seed=7, sigma_c=8, contact=contact, sigma_k=16)
    p = synth_params()
    with lcl.LclSolver(n, n) as solver:
        f, s = solver.solve(p, a, fields=("pressure", "flow_z"))
        f_again, _ = solver.solve(p, a, fields=("pressure",))
        assert np.array_equal(f["pressure"], f_again["pressure"])          # deterministic reductions
        p.rel_tol = 1e-12
        g, s2 = solver.solve(p, a, fields=("pressure", "flow_z"))
    assert s["converged"] == 1 and s["rel_residual"] <= 1e-10 and s["rel_residual_scaled"] <= 1e-10
    main = anchored_open_mask(a)
    assert np.abs(f["pressure"] - g["pressure"])[main].max() / 1000.0 < 1e-6
    assert np.abs(f["flow_z"] - g["flow_z"]).max() <= 1e-6 * np.abs(g["flow_z"]).max()
    # discrete maximum principle where f64 determines the pressure (floating pockets and the contact cells
    # slaved to them are left where the early iterates put them, DESIGN.md "Floating pockets")
    assert f["pressure"][main].min() >= -1e-6 and f["pressure"][main].max() <= 1000.0 + 1e-6
    _, _, trz, _ = lcl.conductance(synth_params(), a)
    q_in, q_out = _inflow_outflow(f, 1000.0, 0.0, trz)
    assert q_in == pytest.approx(q_out, rel=1e-7)
    assert abs(s["flow"]) == pytest.approx(q_out, rel=1e-6)
    assert s["iterations"] < 25   # 11 and 17 measured: a weaker cycle shows here first


def test_torch_cuda_tensor_io_matches_host_io():
    import torch
    a = synth_aperture(300, 200, seed=11, sigma_c=4, contact=0.1, sigma_k=8)
    p = synth_params()
    f_host, s_host = lcl.lcl_solve(p, a)
    f_dev, s_dev = lcl.lcl_solve(p, torch.from_numpy(a).cuda())
    for k in lcl.FIELD_NAMES:
        assert f_dev[k].is_cuda
        assert np.array_equal(f_dev[k].cpu().numpy(), f_host[k], equal_nan=True)
    assert s_dev["flow"] == s_host["flow"]


def test_bulk_solve_matches_single_solves(monkeypatch):
    """BASELINE config 5 in miniature: (map, outlet pressure) sweep through apm_lcl_bulk_solve, every case solved
    on its own (the sharing of unit solutions is the next test)."""
    monkeypatch.setenv("APM_LCL_BULK_SHARE", "0")
    maps = [synth_aperture(96, 128, seed=k, sigma_c=4, contact=0.1, sigma_k=6) for k in range(3)]
    cases = [(synth_params(outlet_pressure=p_out), m) for m in maps for p_out in (0.0, 250.0, 937.5)]
    res = lcl.bulk_solve(cases, devices=(0,), fields=("pressure", "flow_z"))
    assert len(res) == 9
    for (p, m), (arr, sc) in zip(cases, res):
        f, s = lcl.lcl_solve(p, m, fields=("pressure", "flow_z"))
        assert sc["converged"] == 1 and sc["outlet_pressure"] == p.outlet_pressure and sc["iterations"] > 0
        assert np.array_equal(arr["pressure"], f["pressure"]) and sc["flow"] == s["flow"]
    # affine in the boundary pressures: flow scales with (P_in - P_out) on the same map
    q0, q1 = res[0][1]["flow"], res[1][1]["flow"]
    assert q1 / q0 == pytest.approx(750.0 / 1000.0, rel=1e-6)


def test_bulk_sweep_shares_unit_solutions(monkeypatch):
    from helpers import check_bulk_sweep_shares_unit_solutions
    check_bulk_sweep_shares_unit_solutions(lcl, None, lambda **k: synth_params(**k), monkeypatch, shape=(300, 260))


def _device_vector(solver, name):
    import ctypes as C
    ptr, cnt = C.c_void_p(), C.c_size_t()
    assert solver.lib.apm_lcl_solver_device_field(solver._h, name.encode(), C.byref(ptr), C.byref(cnt)) == 0
    return ptr.value, cnt.value


def _memcpy(dst, src, nbytes, kind):
    try:
        from cuda.bindings import runtime as cudart
    except ImportError:                                    # older cuda-python layout
        from cuda import cudart
    kinds = cudart.cudaMemcpyKind
    (err,) = cudart.cudaMemcpy(dst, src, nbytes, getattr(kinds, kind))
    assert int(err) == 0


def _from_device(ptr, count):
    out = np.empty(count)
    _memcpy(out.ctypes.data, ptr, out.nbytes, "cudaMemcpyDeviceToHost")
    return out


@pytest.mark.parametrize("w_levels", [None, (1, 2)])
def test_preconditioner_is_symmetric_positive_on_device(monkeypatch, w_levels):
    """CG needs M = M^T > 0: <M a, b> == <a, M b> for the cycle the device runs (fused down/up legs,
    hybrid tile smoothers, truncated W-cycle), on a map that spans many tiles with ragged edges."""
    if w_levels:
        monkeypatch.setenv("APM_LCL_W_LO", str(w_levels[0]))
        monkeypatch.setenv("APM_LCL_W_HI", str(w_levels[1]))
    nz, nx = 1100, 900
    a = synth_aperture(nz, nx, seed=5, sigma_c=8, contact=0.2, sigma_k=16)
    with lcl.LclSolver(nx, nz) as solver:
        solver.solve(synth_params(), a, fields=("pressure",))
        tx = _from_device(*_device_vector(solver, "tx")).reshape(nz, nx)
        tz = _from_device(*_device_vector(solver, "tz")).reshape(nz, nx)
        txl, tzb = _from_device(*_device_vector(solver, "txl")), _from_device(*_device_vector(solver, "tzb"))
        diag = (np.concatenate([txl[:, None], tx[:, :-1]], axis=1) + tx) + (np.concatenate([tzb[None, :], tz[:-1]], axis=0) + tz)
        (r_ptr, n), (q_ptr, _) = _device_vector(solver, "r"), _device_vector(solver, "q")
        rng = np.random.default_rng(1)
        va, vb = (diag * rng.standard_normal((nz, nx))).ravel(), (diag * rng.standard_normal((nz, nx))).ravel()

        def apply(v):
            v = np.ascontiguousarray(v)
            _memcpy(r_ptr, v.ctypes.data, v.nbytes, "cudaMemcpyHostToDevice")
            solver.time_kernel("vcycle", reps=1)
            return _from_device(q_ptr, n)
        ma, mb = apply(va), apply(vb)
    aa, bb, ab, ba = ma @ va, mb @ vb, ma @ vb, mb @ va
    assert aa > 0 and bb > 0
    assert abs(ab - ba) <= 1e-10 * np.sqrt(aa * bb)
