#!/usr/bin/env python
"""bench.py -- LCL Mcells/s solved (BASELINE.json metric) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 8192] [--impl b200|reference]

A "step" is one pass of the hot path over one synthetic aperture map: aperture map -> conductances
-> multigrid hierarchy -> PCG solve to 1e-10 -> fluxes/derived fields.  N=1 workload: BASELINE
config 3 (8192x8192 log-normal map, ~20 % zero-aperture contact cells), the configuration the metric
is quoted on.  `value` is measured with the aperture already resident in HBM (device pointers in,
device pointers out); `e2e` goes through the same C-ABI call with pinned HOST buffers, so the H2D
copy of the map and the D2H copy of all seven output fields are inside the timed region.
N>1 (torchrun, one rank per GPU): the SAME map is cut into N row strips, one per GPU (BASELINE configs
3/4: strip decomposition, strong scaling); ghost rows and dot products cross the GPUs inside the solver's
kernels (peer stores over NVLink), torch.distributed only carries the IPC handles and the timing.
`--mode bulk` measures the other multi-GPU use instead: independent maps, one per GPU (config 5, weak).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "LCL Mcells/s solved to 1e-10 residual"
UNIT = "Mcells/s"
FIELDS = ("pressure", "flow_x", "flow_z", "flow_m", "aperture", "reynolds", "residual")


def synth_params():
    from apmapflow_b200._lib import Params
    # SURVEY.md 8(d) common .inp values
    return Params.defaults(inlet_pressure=1000.0, outlet_pressure=0.0, density=1000.0, viscosity=1.0e-3,
                           voxel_size=10.0e-6, avg_factor=1.0, roughness=0.0, high_mask=100.0,
                           low_mask=1e-4, outlet_side="TOP")


def synth_map_torch(n, seed, device, contact=0.2, sigma_c=8.0, sigma_k=16.0):
    """CPU generator seeded 20240601+k, N(0,1) field, separable periodic Gaussian
    smoothing (done with torch.fft on the device: input generation is not timed), a = exp(ln5+0.5Z)
    clipped at 100 voxels, lowest `contact` fraction of an independent smoothed field -> 0."""
    import math
    import torch
    g = torch.Generator(device="cpu").manual_seed(20240601 + seed)

    def field(sigma):
        z = torch.randn((n, n), generator=g, dtype=torch.float64).to(device)
        r = int(4 * sigma)
        k = torch.zeros(n, dtype=torch.float64, device=device)
        idx = torch.arange(-r, r + 1, device=device)
        k.index_add_(0, idx % n, torch.exp(-0.5 * (idx.double() / sigma) ** 2))
        k /= k.sum()
        kf = torch.fft.rfft(k)
        z = torch.fft.irfft(torch.fft.rfft(z, dim=1) * kf[None, :], n=n, dim=1)
        kf0 = torch.fft.fft(k)
        z = torch.fft.ifft(torch.fft.fft(z, dim=0) * kf0[:, None], dim=0).real
        return (z - z.mean()) / z.std()

    a = torch.clamp(torch.exp(math.log(5.0) + 0.5 * field(sigma_c)), max=100.0)
    if contact > 0:
        c = field(sigma_k)
        kth = max(1, int(contact * n * n))
        thr = torch.kthvalue(c.flatten(), kth).values
        a = torch.where(c < thr, torch.zeros_like(a), a)
    return a.contiguous()


class ClockSampler(object):
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md's
    clocks line).  NVML is polled in-process every 10 ms -- a timed region of a few solves lasts a few
    hundred ms, less than it takes an nvidia-smi process to start; nvidia-smi -lms is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index   # rows: (sm MHz, max MHz, set of reasons)
        self.nvml, self.handle, self.thread, self.stop_flag = None, None, None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            dev = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [d for d in dev.split(",") if d.strip()]
            phys = int(ids[index]) if ids and all(d.strip().isdigit() for d in ids) and index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n, h = self.nvml, self.handle
        bits = ((n.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                (n.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (n.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap"))
        mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
        while not self.stop_flag.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((sm, mx, {name for bit, name in bits if mask & bit}))
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((float(r[0]), float(r[1]),
                                  {name for name, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")}))
            except (ValueError, IndexError):
                continue

    def start(self):
        if self.nvml:
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(r[0] for r in self.rows)
        reasons = set().union(*[r[2] for r in self.rows]) if self.rows else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.rows[-1][1] if self.rows else None,
                "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi", "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(sample_n, seed=0, contact=0.2):
    """The oracle's direct D4 solve (restatement of the reference Fortran) on a bounded sample of
    the same workload recipe, one core."""
    import numpy as np
    from oracle import oracle_binding as ob
    import torch
    a = synth_map_torch(sample_n, seed, "cpu", contact=contact).numpy()
    t0 = time.time()
    _, sc = ob.solve(synth_params(), a, want=("pressure",))
    wall = time.time() - t0
    return {"value": sample_n * sample_n / wall / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%dx%d map of the same recipe, whole oracle path %.2f s (D4 solve %.2f s); D4 work grows as n^4"
                      % (sample_n, sample_n, wall, sc["solve_seconds"])}, a, sc


_REF_MAPS = {}


def _ref_prepare(args):
    """Generate (untimed) and keep the worker's map: pool workers are pinned to one task each."""
    n, seed = args
    import torch
    torch.set_num_threads(1)
    from oracle import oracle_binding as ob
    ob.load()
    _REF_MAPS[(n, seed)] = synth_map_torch(n, seed, "cpu").numpy()
    return os.getpid()


def _ref_worker(args):
    n, seed = args
    from oracle import oracle_binding as ob
    a = _REF_MAPS.get((n, seed))
    if a is None:                      # the task landed on another worker than its prepare step
        a = synth_map_torch(n, seed, "cpu").numpy()
    t0 = time.time()
    ob.solve(synth_params(), a, want=("pressure",))
    return time.time() - t0


def run_reference(args, emit):
    """--impl reference: the reference's CPU algorithm (oracle port; gfortran is absent so oracle/_ref
    cannot exist) on all host cores, one independent map per core like BulkRun would schedule them."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n = args.ref_size
    from oracle import oracle_binding as ob
    ob.build()
    times = []
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            tasks = [(n, 1000 * step + k) for k in range(cores)]
            pool.map(_ref_prepare, tasks, chunksize=1)          # synthetic inputs: not timed
            t0 = time.time()
            per_task = pool.map(_ref_worker, tasks, chunksize=1)
            wall = time.time() - t0
            if step >= args.warmup:
                times.append(max(wall, max(per_task)))
    ms = 1e3 * sum(times) / len(times)
    value = cores * n * n / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if (args.gpus > 1 and args.mode == "striped") else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%d independent %dx%d log-normal maps with ~20%% contact cells per step, one per "
                                   "host core (bounded sample of the 8192x8192 workload: the reference's direct D4 "
                                   "solve needs 0.34 n^3 doubles = 1.5 TB at 8192)" % (cores, n, n)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d x (%dx%d) maps per step" % (cores, n, n)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--contact", type=float, default=0.2)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ref-size", type=int, default=384)
    ap.add_argument("--cpu-sample", type=int, default=448)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="striped", choices=["striped", "bulk"],
                    help="N>1: one map striped over the GPUs (strong scaling) or one independent map per GPU (weak)")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner under
    # torchrun does) are sent to stderr until the line is ready.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)

    if args.impl == "reference":
        return run_reference(args, emit)

    import numpy as np
    import torch
    import torch.distributed as dist
    from apmapflow_b200 import lcl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = 0
    if world > 1:
        from apmapflow_b200 import dist as apm_dist
        numa_cores = apm_dist.bind_to_gpu_numa(local)     # pinned buffers land on the GPU's own NUMA node
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.size
    cells = n * n
    striped = world > 1 and args.mode == "striped"
    # every rank generates the map of seed 20240601 (input generation is not timed).  Striped: each rank keeps
    # the rows it needs.  Bulk: the same map on every rank, so that per-GPU work is fixed (maps of other seeds
    # need 13-16 iterations at 8192^2, which would measure the maps, not the scaling: profiles/README.md)
    a_dev = synth_map_torch(n, 0, dev, contact=args.contact)
    p = synth_params()
    if striped:
        solver = lcl.StripedLclSolver(n, n, device=local)
        a_dev = a_dev[solver.a_lo:solver.a_hi].clone()
        rows = solver.z_hi - solver.z_lo
        torch.cuda.empty_cache()
    else:
        solver = lcl.LclSolver(n, n, device=local)
        rows = n
    outs = {k: torch.empty((rows, n), dtype=torch.float64, device=dev) for k in FIELDS}
    units = cells if striped else world * cells          # cells all ranks solve per step

    def step_device():
        fo = lcl.Fields()
        for k in FIELDS:
            setattr(fo, k, outs[k].data_ptr())
        sc = lcl.Scalars()
        import ctypes as C
        rc = solver.lib.apm_lcl_solver_solve(solver._h, C.byref(p), a_dev.data_ptr(), C.byref(fo), C.byref(sc))
        if rc:
            raise RuntimeError(solver.lib.apm_lcl_last_error().decode())
        return sc

    # ---------------------------------------------------------------- device-resident timing
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        sc = step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib_ms, launches, iters = 0.0, 0, 0
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        sc = step_device()          # the library times its own stream with CUDA events (total_ms)
        lib_ms += sc.total_ms
        launches += int(solver.lib.apm_lcl_last_launch_count())
        iters += sc.iterations
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    # device time: CUDA events recorded by the library on the stream its kernels run on (first kernel
    # -> last output copy of every step); the host wall clock around the same steps is the cross-check
    ms_step = lib_ms / args.steps
    t = torch.tensor([ms_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    value = units / (ms_step * 1e-3) / 1e6
    last = {k: getattr(sc, k) for k, _ in lcl.Scalars._fields_}

    # ---------------------------------------------------------------- end to end, host buffers
    a_host = a_dev.cpu().pin_memory()
    host_out = {k: torch.empty((rows, n), dtype=torch.float64).pin_memory() for k in FIELDS}

    def step_host():
        import ctypes as C
        fo = lcl.Fields()
        for k in FIELDS:
            setattr(fo, k, host_out[k].data_ptr())
        s2 = lcl.Scalars()
        rc = solver.lib.apm_lcl_solver_solve(solver._h, C.byref(p), a_host.data_ptr(), C.byref(fo), C.byref(s2))
        if rc:
            raise RuntimeError(solver.lib.apm_lcl_last_error().decode())
        return s2

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    # bytes crossing PCIe per step, summed over the ranks (a strip copies its rows plus one either side in, its
    # own rows of the seven fields out)
    t = torch.tensor([8.0 * a_host.numel(), 8.0 * rows * n * len(FIELDS)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    e2e = {"value": units / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(t[0].item()), "d2h_bytes_per_step": int(t[1].item())}

    def finish():
        if striped:
            solver.close()
        if world > 1:
            dist.destroy_process_group()

    if rank != 0 and not striped:
        return finish()

    # ---------------------------------------------------------------- roofline of the kernels
    # (striped: every rank times the kernels of ITS strip -- they are collective, the dot products meet across
    # the GPUs inside them -- and rank 0 reports its own)
    peak, peak_src = measured_peak()
    own_cells = rows * n
    step_device()
    kern = {}
    # launches per PCG iteration of each fine-level kernel (+ the level-1 kernels of the V-cycle)
    per_iter = {"spmv": 1, "update": 1, "direction": 1, "down5": 1, "up5": 1,
                "smooth9_l1": 1, "residual9_l1": 1, "restrict1": 1, "smooth9_l1_up": 1, "vcycle": 1}
    for name in per_iter:
        ms, nbytes = solver.time_kernel(name, reps=20)
        step_device() if name != list(per_iter)[-1] else None
        kern[name] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak,
                      "bytes_per_cell": nbytes / own_cells, "launches_per_iter": per_iter[name]}
    dom = max((k for k in kern if k != "vcycle"), key=lambda k: kern[k]["ms"] * kern[k]["launches_per_iter"])
    if rank != 0:
        return finish()
    traffic = None
    # ncu capture of these kernels, final code of round 2 (tools/r02_final_profile.sh + tools/ncu_raw_to_json.py)
    tpath = os.path.join(ROOT, "profiles", "r02_ncu_traffic_final.json")
    if os.path.exists(tpath) and world == 1:
        tj = json.load(open(tpath))
        if tj.get("size") == n and dom in tj["kernels"]:
            traffic = tj["kernels"][dom]["dram_read_bytes"] + tj["kernels"][dom]["dram_write_bytes"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": traffic,
                "algorithmic_bytes": kern[dom]["bytes_per_cell"] * own_cells, "peak_source": peak_src, "kernels": kern}

    cpu = None
    if not args.no_cpu_baseline and world == 1:     # the CPU baseline is timed on rank 0 at N = 1 only
        cpu, _, _ = cpu_baseline(args.cpu_sample)

    how = ("ONE map cut into %d row strips, one per GPU (ghost rows and dot products by peer stores over NVLink)" % world
           if striped else "one map per GPU (the same map, seed 20240601, on every rank)")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if striped else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%dx%d synthetic log-normal aperture map (sigma_c=8 cells), %.0f%% zero-aperture "
                                   "contact cells (sigma=16), constant-pressure inlet/outlet (1000 -> 0 Pa, TOP), %s"
                                   % (n, n, 100 * args.contact, how),
                       "cells_per_gpu": own_cells, "rel_tol": 1e-10, "pcg_iterations": iters / args.steps,
                       "host_cores_bound_to_gpu_numa_node": numa_cores,
                       "l2": "inputs larger than L2 (%.0f MB per field vs 126 MB)" % (8 * cells / 1e6),
                       "host_wall_ms_per_step": wall_ms / args.steps,
                       "setup_ms": last["setup_ms"], "solve_ms": last["solve_ms"],
                       "rel_residual": last["rel_residual"], "rel_residual_scaled": last["rel_residual_scaled"]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    emit(line)
    finish()


if __name__ == "__main__":
    main()
