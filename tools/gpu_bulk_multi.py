"""GPU helper (run under gpurun --gpus N): BASELINE config 5 -- 16 maps x 16 outlet pressures at 1024^2, the 256 cases
sharded over all GPUs of the box (one host thread per device, apm_lcl_bulk_solve on each: independent cases, no
collective), host numpy buffers in and out, wall clock around the whole sweep."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from helpers import synth_aperture
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params, TOP

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nmaps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ngpu = torch.cuda.device_count()
maps = [synth_aperture(n, n, seed=k, sigma_c=8, contact=0.2, sigma_k=16) for k in range(nmaps)]
cases = [(Params.defaults(inlet_pressure=1000.0, outlet_pressure=62.5 * j, density=1000.0, viscosity=1e-3, voxel_size=1e-5,
                          avg_factor=1.0, roughness=0.0, high_mask=100.0, low_mask=1e-4, outlet_side=TOP), m)
         for m in maps for j in range(16)]
for g in sorted({1, 2, 4, ngpu} & set(range(1, ngpu + 1))):
    devs = tuple(range(g))
    lcl.bulk_solve(cases[:4 * g], devices=devs)          # warm-up (contexts, module load)
    t0 = time.time()
    res = lcl.bulk_solve(cases, devices=devs, fields=("pressure",))
    dt = time.time() - t0
    ok = all(s["converged"] == 1 for _, s in res)
    print("config 5: %d cases of %d^2 on %d GPU(s): wall %.3f s = %.2f ms/case = %.0f Mcells/s, all converged=%s"
          % (len(cases), n, g, dt, 1e3 * dt / len(cases), len(cases) * n * n / dt / 1e6, ok), flush=True)
