"""GPU debugging helper (run under gpurun): solves a ladder of synthetic maps and prints stats."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from apmapflow_b200 import lcl
import bench

sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [512, 1024, 2048, 4096]
contacts = [float(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.0, 0.2]
for n in sizes:
    for c in contacts:
        a = bench.synth_map_torch(n, 0, "cuda", contact=c)
        p = bench.synth_params()
        try:
            f, s = lcl.lcl_solve(p, a, fields=("pressure",))
            print("n=%d contact=%.1f iters=%d conv=%d res=%.2e scaled=%.2e setup=%.2fms solve=%.2fms total=%.2fms launches=%d Mcells/s=%.1f"
                  % (n, c, s["iterations"], s["converged"], s["rel_residual"], s["rel_residual_scaled"], s["setup_ms"],
                     s["solve_ms"], s["total_ms"], s["gpu_launches"], n * n / s["total_ms"] / 1e3), flush=True)
        except Exception as e:
            print("n=%d contact=%.1f FAILED: %s %s" % (n, c, e, getattr(e, "scalars", None)), flush=True)
