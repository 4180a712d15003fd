"""GPU helper (run under gpurun): BASELINE config 5 -- 16 maps x 16 outlet pressures at 1024^2 through
apm_lcl_bulk_solve -- timed for several numbers of concurrent solvers per device, every case solved on its own
(APM_LCL_BULK_SHARE=0: the honest 256-solve number), then with the cases of one map sharing their unit solution
(the library's default: 16 solves + 240 affine maps and flux passes; reported separately)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from helpers import synth_aperture
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params, TOP

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nmaps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
maps = [synth_aperture(n, n, seed=k, sigma_c=8, contact=0.2, sigma_k=16) for k in range(nmaps)]
cases = []
for m in maps:
    for j in range(16):
        cases.append((Params.defaults(inlet_pressure=1000.0, outlet_pressure=62.5 * j, density=1000.0, viscosity=1e-3,
                                      voxel_size=1e-5, avg_factor=1.0, roughness=0.0, high_mask=100.0, low_mask=1e-4,
                                      outlet_side=TOP), m))
lcl.bulk_solve(cases[:8], devices=(0,))          # warm-up (context, module load)
for share, stream_counts in (("0", (1, 2, 4, 8)), ("1", (1, 4))):
    os.environ["APM_LCL_BULK_SHARE"] = share
    for streams in stream_counts:
        os.environ["APM_LCL_BULK_STREAMS"] = str(streams)
        dt = 1e30
        for _ in range(3):             # best of three: the wall clock of this host-fed loop is noisy on a shared box
            t0 = time.time()
            res = lcl.bulk_solve(cases, devices=(0,), fields=("pressure",))
            dt = min(dt, time.time() - t0)
        ok = all(s["converged"] == 1 for _, s in res)
        solved = sum(s["iterations"] > 0 for _, s in res)
        print("n=%d cases=%d (PCG solves: %d) share=%s streams=%d wall=%.3f s  %.2f ms/case  %.1f Mcells/s  all converged=%s" %
              (n, len(cases), solved, share, streams, dt, 1e3 * dt / len(cases), len(cases) * n * n / dt / 1e6, ok), flush=True)
