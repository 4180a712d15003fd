"""GPU helper: stage timings of apm_lcl_percentiles / apm_lcl_map_stats (APM_LCL_STATS_TIMING=1) on the bench map."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from apmapflow_b200 import lcl
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = bench.synth_map_torch(n, 0, torch.device("cuda", 0))
pct = [0, 1, 5, 10, 15, 20, 25, 30, 40, 50, 60, 70, 75, 80, 85, 90, 95, 99, 100]
lcl.percentiles(a, pct); lcl.map_stats(a); torch.cuda.synchronize()
for tag, arr in (("device tensor", a), ("host array", a.cpu().numpy())):
    print("==", tag, flush=True)
    t0 = time.time(); p = lcl.percentiles(arr, pct); t1 = time.time(); s = lcl.map_stats(arr); t2 = time.time()
    print("%s: 19 percentiles %.1f ms, map_stats %.1f ms" % (tag, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
