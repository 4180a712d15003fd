"""Times apm_lcl_run_file (text map in, CSV/VTK/stat files out) on a synthetic map under gpurun."""
import os, sys, time, tempfile, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, pandas as pd
from apmapflow_b200 import _lib
from helpers import synth_aperture
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
tmp = tempfile.mkdtemp()
a = synth_aperture(n, n, seed=1, sigma_c=8, contact=0.2, sigma_k=16)
t0 = time.time(); pd.DataFrame(a).to_csv(os.path.join(tmp, "map.txt"), sep="\t", header=False, index=False, float_format="%.4f"); print("map written %.1fs" % (time.time() - t0))
inp = ("APER-MAP: %s/map.txt\nSUMMARY-FILE: %s/log.txt\nSTAT-FILE: %s/stat.csv\nAPER-FILE: %s/aper.csv\nFLOW-FILE: %s/flow.csv\n"
       "PRESS-FILE: %s/pres.csv\nVTK-FILE: %s/data.vtk\nOVERWRITE\nINLET-PRESS: 1000 PA\nOUTLET-PRESS: 0 PA\nOUTLET-SIDE: TOP\n"
       "FLUID-DENSITY: 1000 KG/M^3\nFLUID-VISCOSITY: 1.0 CP\nMAXIMUM MAP DIMENSION: 9999\nMAP AVERAGING FACTOR: 1\nVOXEL SIZE: 10 MICRONS\n"
       "HIGH-MASK: 100\nLOW-MASK: 0.0001\n"
       "CALCULATE PERCENTILES: 0,1,5,10,15,20,25,30,40,50,60,70,75,80,85,90,95,99,100\n") % ((tmp,) * 7)
open(os.path.join(tmp, "case.inp"), "w").write(inp)
lib = _lib.load(); buf = C.create_string_buffer(1 << 22)
t0 = time.time(); rc = lib.apm_lcl_run_file(os.path.join(tmp, "case.inp").encode(), 0, buf, len(buf)); dt = time.time() - t0
out = buf.value.decode()
print("rc", rc, "run_file wall %.2fs for %dx%d" % (dt, n, n))
print([l for l in out.split("\n") if "REQUIRED TIME" in l or "MULTIGRID" in l or "NET FLOW" in l])
sizes = {f: os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp)}
print("output bytes", sum(sizes.values()) - sizes["map.txt"] - sizes["case.inp"], "input bytes", sizes["map.txt"])
