"""Debug aid (GPU box): the striped solve with every rank on GPU 0, in the ways a single-GPU box allows.
    python tools/gpu_striped_debug.py threads|procs [world] [nz] [nx]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from helpers import synth_aperture
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params

mode = sys.argv[1]
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nz = int(sys.argv[3]) if len(sys.argv) > 3 else 512
nx = int(sys.argv[4]) if len(sys.argv) > 4 else 384
fields = tuple(os.environ.get("DBG_FIELDS", "pressure").split(",")) if os.environ.get("DBG_FIELDS", "pressure") else ()
os.environ.setdefault("APM_LCL_STRIPE_MIN_CELLS", "3000")
os.environ.setdefault("APM_LCL_PEER_TIMEOUT_MS", "4000")
a = synth_aperture(nz, nx, seed=nz + world, sigma_c=6, contact=0.15, sigma_k=10)
p = Params.defaults(inlet_pressure=1000.0, outlet_pressure=0.0, density=1000.0, viscosity=1e-3, voxel_size=1e-5,
                    avg_factor=1.0, high_mask=100.0, low_mask=1e-4, outlet_side="TOP")
if mode == "threads":
    t0 = time.time()
    try:
        with lcl.StripedGroup(nx, nz, [0] * world) as g:
            print("rows", g.rows, flush=True)
            whole, s = g.solve(p, a, fields=fields)
            print("OK threads: iters", s["iterations"], "flow", s["flow"], "total_ms", s["total_ms"], "wall", time.time() - t0, flush=True)
            f1, s1 = lcl.lcl_solve(p, a, fields=("pressure",))
            print("single: iters", s1["iterations"], "flow", s1["flow"], "total_ms", s1["total_ms"])
            if "pressure" in whole:
                print("max |dp|", np.abs(whole["pressure"] - f1["pressure"]).max())
    except Exception as e:
        print("FAILED threads:", e, flush=True)
        for sc in getattr(e, "scalars", []) or []:
            print("   ", sc)
else:
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
    try:
        with lcl.StripedLclSolver(nx, nz, device=0) as solver:
            f, s = solver.solve(p, a, fields=fields)
            print("OK procs rank", rank, "iters", s["iterations"], "flow", s["flow"], "total_ms", s["total_ms"], flush=True)
    except Exception as e:
        print("FAILED procs rank", rank, e, flush=True)
    dist.destroy_process_group()
