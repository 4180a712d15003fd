"""ncu target: one solve of the bench map, then each hot kernel launched alone inside an NVTX range "roofline"
(apm_lcl_solver_time_kernel: 3 warm-up launches + 1), so that
    ncu --set full --nvtx --nvtx-include "roofline/" ... python tools/ncu_roofline.py [size] [rows]
captures exactly those launches.  rows < size profiles a strip-shaped map (rows x size), the kernels' view of one
rank of a striped solve (e.g. 2048 x 16384 = the 16384^2 / 8 strip)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200")):
    sys.path.insert(0, p)
import torch
import bench
from apmapflow_b200 import lcl

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rows = int(sys.argv[2]) if len(sys.argv) > 2 else size
dev = torch.device("cuda", 0)
a = bench.synth_map_torch(size, 0, dev)[:rows].contiguous()
p = bench.synth_params()
solver = lcl.LclSolver(size, rows, device=0)
f, s = solver.solve(p, a, fields=("pressure",))
print("solve:", rows, "x", size, s["iterations"], "iterations", round(s["total_ms"], 2), "ms", flush=True)
names = ["spmv", "update", "direction", "down5", "up5", "smooth9_l1", "residual9_l1", "restrict1", "smooth9_l1_up"]
for name in names:
    solver.solve(p, a, fields=("pressure",))
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("roofline")
    ms, nbytes = solver.time_kernel(name, reps=1)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print(name, round(ms, 4), "ms", round(nbytes / ms / 1e6, 1), "GB/s", flush=True)
