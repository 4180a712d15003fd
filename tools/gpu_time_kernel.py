"""GPU helper: CUDA-event time of single kernels on the solved bench map (apm_lcl_solver_time_kernel).
    [APM_LCL_LIB=<other build>] python tools/gpu_time_kernel.py 8192 flux conductance"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402
from apmapflow_b200 import lcl  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
names = sys.argv[2:] or ["flux"]
a = bench.synth_map_torch(size, 0, torch.device("cuda", 0)).cpu().numpy()      # host map: "conductance" needs the staged copy
solver = lcl.LclSolver(size, size, device=0)
for name in names:
    f, s = solver.solve(bench.synth_params(), a, fields=("pressure",))
    ms, nbytes = solver.time_kernel(name, reps=20)
    print("%s: %.4f ms, %.0f GB/s (%.1f B/cell), solve %.2f ms setup %.2f ms" %
          (name, ms, nbytes / ms / 1e6, nbytes / size / size, s["total_ms"], s["setup_ms"]), flush=True)
