"""One map striped over several ranks on the GPU (-m gpu), through libapmlcl.so: ghost rows and dot products
cross the cuts by peer stores and flags inside the kernels (csrc/dist.cuh, csrc/peer.cu).  The result must be
the single-GPU result: the strips run the same iteration, only the order of the dot-product sums differs.

Two ways to run the ranks: threads of this process (lcl.StripedGroup: peers addressed directly; with every
rank on GPU 0 this also runs on a single-GPU box) and one process per rank over CUDA IPC (what torchrun /
bench.py --gpus N do; needs >= 2 GPUs)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from helpers import anchored_open_mask, synth_aperture
from oracle import oracle_binding as ob
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params, LEFT, TOP, BOTTOM

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synth_params(**kw):
    base = dict(inlet_pressure=1000.0, outlet_pressure=0.0, density=1000.0, viscosity=1.0e-3,
                voxel_size=10.0e-6, avg_factor=1.0, roughness=0.0, high_mask=100.0, low_mask=1e-4,
                outlet_side=TOP)
    base.update(kw)
    return Params.defaults(**base)


def gpu_count():
    import torch
    return torch.cuda.device_count()


def compare_with_single(whole, s, f1, s1):
    assert s["converged"] == 1
    assert (s["zeros"], s["masked"], s["min_aper"], s["max_aper"]) == (s1["zeros"], s1["masked"], s1["min_aper"], s1["max_aper"])
    assert abs(s["iterations"] - s1["iterations"]) <= 1
    assert s["flow"] == pytest.approx(s1["flow"], rel=1e-9)
    assert np.abs(whole["pressure"] - f1["pressure"]).max() <= 2e-7 * np.abs(f1["pressure"]).max()
    assert np.array_equal(whole["aperture"], f1["aperture"])
    for k in ("flow_x", "flow_z", "flow_m"):
        assert np.abs(whole[k] - f1[k]).max() <= 1e-6 * np.abs(f1["flow_m"]).max()
    assert np.nanmax(whole["residual"]) == 1.0


@pytest.mark.parametrize("world,shape,side", [(2, (512, 384), TOP), (3, (700, 300), LEFT), (4, (1024, 256), BOTTOM)])
def test_strips_on_one_gpu_match_the_single_solve_and_the_oracle_recipe(monkeypatch, world, shape, side):
    """Every rank on GPU 0 (threads, shared device): the whole striped control flow on a single-GPU box."""
    monkeypatch.setenv("APM_LCL_STRIPE_MIN_CELLS", "3000")
    nz, nx = shape
    a = synth_aperture(nz, nx, seed=nz + world, sigma_c=6, contact=0.15, sigma_k=10)
    p = synth_params(outlet_side=side)
    with lcl.StripedGroup(nx, nz, [0] * world) as g:
        assert [r[:2] for r in g.rows][0][0] == 0 and g.rows[-1][1] == nz and len(g.rows) == world
        assert all(g.rows[k][1] == g.rows[k + 1][0] for k in range(world - 1))      # the strips tile the map
        whole, s = g.solve(p, a)
        again, s_again = g.solve(p, a)
        assert np.array_equal(whole["pressure"], again["pressure"])                 # deterministic, also across ranks
        assert len({sc["flow"] for sc in s["per_rank"]}) == 1
    f1, s1 = lcl.lcl_solve(p, a)
    compare_with_single(whole, s, f1, s1)


def test_strips_vs_oracle_small(monkeypatch):
    monkeypatch.setenv("APM_LCL_STRIPE_MIN_CELLS", "1000")
    a = synth_aperture(256, 200, seed=9, sigma_c=6, contact=0.15, sigma_k=10)
    p = synth_params()
    whole, s = lcl.lcl_solve_striped(p, a, [0, 0])
    o, so = ob.solve(p, a)
    mask = anchored_open_mask(a)
    assert (s["zeros"], s["masked"]) == (so["zeros"], so["masked"])
    assert (np.abs(whole["pressure"] - o["pressure"]) / 1000.0)[mask].max() < 1e-6
    assert s["flow"] == pytest.approx(so["flow"], rel=1e-6)
    assert np.array_equal(whole["aperture"], o["width"])


def test_rate_bc_and_back_pressure_striped(monkeypatch):
    monkeypatch.setenv("APM_LCL_STRIPE_MIN_CELLS", "3000")
    a = synth_aperture(384, 256, seed=2, sigma_c=6)
    p = synth_params(outlet_side=BOTTOM, inlet_pressure=-1.0, outlet_pressure=2.0e6, inlet_rate=1e-9)
    whole, s = lcl.lcl_solve_striped(p, a, [0, 0, 0])
    f1, s1 = lcl.lcl_solve(p, a)
    assert s["outlet_rate"] == pytest.approx(1e-9, rel=1e-6)
    assert s["inlet_pressure"] == pytest.approx(s1["inlet_pressure"], rel=1e-9)
    assert np.abs(whole["flow_z"] - f1["flow_z"]).max() <= 1e-6 * np.abs(f1["flow_z"]).max()


WORKER = r"""
import os, sys, pickle
sys.path[:0] = [{root!r}, os.path.join({root!r}, 'netl-ap-map-flow_b200'), os.path.join({root!r}, 'tests')]
import numpy as np, torch, torch.distributed as dist
from helpers import synth_aperture
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dev = rank % torch.cuda.device_count()        # fewer GPUs than ranks: the ranks share them (time sliced)
torch.cuda.set_device(dev)
dist.init_process_group('gloo')
nz, nx = {nz}, {nx}
a = synth_aperture(nz, nx, seed=5, sigma_c=8, contact=0.2, sigma_k=16)
p = Params.defaults(inlet_pressure=1000.0, outlet_pressure=0.0, density=1000.0, viscosity=1e-3, voxel_size=1e-5,
                    avg_factor=1.0, high_mask=100.0, low_mask=1e-4, outlet_side='TOP')
with lcl.StripedLclSolver(nx, nz, device=dev) as solver:
    f, s = solver.solve(p, torch.from_numpy(a[solver.a_lo:solver.a_hi]).cuda(dev))
    whole = solver.gather(f)
if rank == 0:
    f1, s1 = lcl.lcl_solve(p, a, device=0)
    pickle.dump((whole, s, f1, s1), open({out!r}, 'wb'))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_process_per_gpu_over_cuda_ipc(tmp_path, world):
    """torchrun-style: one process per rank, slabs mapped through CUDA IPC, stores over NVLink.  On a box with
    fewer GPUs than ranks the processes share GPUs (their kernels are time sliced, every wait burns a slice:
    slow but the same code path), so two ranks are also exercised on a single-GPU box."""
    if gpu_count() < world and world > 2:
        pytest.skip("needs %d GPUs" % world)
    nz, nx = (2048, 1536) if gpu_count() >= world else (512, 384)
    script, out = tmp_path / "worker.py", tmp_path / "out.pkl"
    script.write_text(WORKER.format(root=ROOT, nz=nz, nx=nx, out=str(out)))
    port = 29500 + os.getpid() % 2000
    env = dict(os.environ, APM_LCL_STRIPE_MIN_CELLS="20000" if gpu_count() >= world else "3000")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                    "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], check=True, env=env,
                   timeout=600)
    whole, s, f1, s1 = pickle.load(open(out, "rb"))
    compare_with_single(whole, s, f1, s1)


def test_c_level_one_shot_striped_solve(monkeypatch):
    """apm_lcl_solve_striped: whole-map host buffers in and out, strips driven by the library's own threads."""
    import ctypes as C
    from apmapflow_b200 import _lib
    from apmapflow_b200._lib import Fields, Scalars
    monkeypatch.setenv("APM_LCL_STRIPE_MIN_CELLS", "3000")
    lib = _lib.load()
    nz, nx = 320, 272
    a = synth_aperture(nz, nx, seed=21, sigma_c=6, contact=0.1, sigma_k=10)
    p = synth_params()
    pres, qz = np.empty((nz, nx)), np.empty((nz, nx))
    fo, sc = Fields(pressure=pres.ctypes.data, flow_z=qz.ctypes.data), Scalars()
    devs = (C.c_int * 3)(0, 0, 0)
    assert lib.apm_lcl_solve_striped(C.byref(p), a.ctypes.data, nx, nz, 3, devs, C.byref(fo), C.byref(sc)) == 0, _lib.last_error()
    f1, s1 = lcl.lcl_solve(p, a, fields=("pressure", "flow_z"))
    assert sc.converged == 1 and sc.flow == pytest.approx(s1["flow"], rel=1e-9)
    assert np.abs(pres - f1["pressure"]).max() <= 2e-7 * 1000.0
    assert np.abs(qz - f1["flow_z"]).max() <= 1e-6 * np.abs(f1["flow_z"]).max()
    assert lib.apm_lcl_last_launch_count() > 0


def test_run_model_with_devices_writes_the_same_files(tmp_path, monkeypatch):
    """The file-level drop-in with the map striped over several 'devices': run_model(devices=[...]) must write
    what the single-GPU run writes (the formatted digits are the same solution)."""
    from apmapflow_b200.run_model import InputFile, run_model
    monkeypatch.setenv("APM_LCL_STRIPE_MIN_CELLS", "2000")
    golden = os.path.join(ROOT, "tests", "golden")
    outs = {}
    for tag, devices in (("one", None), ("striped", [0, 0])):
        work = tmp_path / tag
        work.mkdir()
        inp = InputFile(os.path.join(golden, "test-model-inputs.txt"))
        for key in ("APER-MAP", "SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE"):
            inp[key].commented_out = False
        for key in ("SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE"):
            inp.filename_formats[key] = os.path.join(str(work), key + ".out")
        inp.filename_formats["input_file"] = os.path.join(str(work), "case.inp")
        inp["APER-MAP"] = os.path.join(golden, "Fracture1ApertureMap-10avg.txt")
        proc = run_model(inp, synchronous=True, devices=devices)
        assert proc.poll() == 0, proc.stdout_content
        outs[tag] = work
    for name in ("PRESS-FILE.out", "FLOW-FILE-z.out", "FLOW-FILE-x.out", "APER-FILE.out"):
        a = np.loadtxt(outs["one"] / name, delimiter=",")
        b = np.loadtxt(outs["striped"] / name, delimiter=",")
        assert np.abs(a - b).max() <= 1e-6 * np.abs(a).max(), name


def test_config3_full_size_8192_strips_equal_the_single_solve():
    """BASELINE config 3 at full size: the 8192^2 map with 20 % contact cells, cut into 2 and into 4 row strips (all on
    GPU 0: the striped code path, not the scaling), against the single-GPU solve of the same map; plus the
    size-independent properties: both residual norms met, inflow == outflow, determinism."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    n = 8192
    dev = torch.device("cuda", 0)
    a = bench.synth_map_torch(n, 0, dev)
    p = bench.synth_params()
    with lcl.LclSolver(n, n) as solver:
        f1, s1 = solver.solve(p, a, fields=("pressure", "flow_z"))
        f1b, _ = solver.solve(p, a, fields=("pressure",))
        assert torch.equal(f1["pressure"], f1b["pressure"])                        # deterministic
        tz, _ = _dev_array(solver, "tz", n * n)
        tzb, _ = _dev_array(solver, "tzb", n)
        q_in = float((tzb * (1000.0 - f1["pressure"][0])).sum())
        q_out = float((tz.view(n, n)[-1] * (f1["pressure"][-1] - 0.0)).sum())
    assert s1["converged"] == 1 and s1["rel_residual"] <= 1e-10 and s1["rel_residual_scaled"] <= 1e-10
    assert q_in == pytest.approx(q_out, rel=1e-7) and abs(s1["flow"]) == pytest.approx(q_out, rel=1e-6)
    assert s1["iterations"] <= 16                                                   # 13 measured
    a_host = a.cpu().numpy()
    ref = f1["pressure"].cpu().numpy()
    del f1, f1b, a
    torch.cuda.empty_cache()
    for world in (2, 4):
        whole, s = lcl.lcl_solve_striped(p, a_host, [0] * world, fields=("pressure",))
        assert s["converged"] == 1 and abs(s["iterations"] - s1["iterations"]) <= 1
        assert s["flow"] == pytest.approx(s1["flow"], rel=1e-9)
        assert np.abs(whole["pressure"] - ref).max() <= 2e-7 * 1000.0
        assert (s["zeros"], s["masked"]) == (s1["zeros"], s1["masked"])


def _dev_array(solver, name, count):
    import ctypes as C
    import torch
    ptr, cnt = C.c_void_p(), C.c_size_t()
    assert solver.lib.apm_lcl_solver_device_field(solver._h, name.encode(), C.byref(ptr), C.byref(cnt)) == 0
    assert cnt.value == count
    out = torch.empty(count, dtype=torch.float64, device="cuda:0")
    from cuda.bindings import runtime as cudart
    (err,) = cudart.cudaMemcpy(out.data_ptr(), ptr.value, count * 8, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert int(err) == 0
    return out, cnt.value
