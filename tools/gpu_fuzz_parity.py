"""Randomised parity run against the oracle (run under gpurun; `--sim` drives the CPU emulation instead, slowly):
random shapes, outlet sides, boundary-condition modes, contact fractions, roughness and masks, each case checked
with the assertions of tests/test_gpu_parity.py::check_against_oracle.  Prints one line per failure with the seed
that reproduces it and a summary; exit status 1 if anything failed.

Maps without an open path from the inlet edge to the outlet edge are skipped: a blocked fracture carries only the
LOW-MASK leak (1e-14 of an open cell's conductance), local fluxes and the flow-rate boundary condition are then
decided by rounding -- the reference's own f64 and quad builds (the oracle's two variants) disagree by 5-15 % there,
and an iterative solve to 1e-10 leaves flux noise of 1e-10 T_open dP, far above the leak.

    python tools/gpu_fuzz_parity.py [--seconds 60] [--seed 0] [--max-size 160] [--sim]
"""
import argparse
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from helpers import anchored_open_mask, synth_aperture  # noqa: E402
from apmapflow_b200._lib import LEFT, RIGHT, TOP, BOTTOM  # noqa: E402


def draw_case(rng, max_size):
    small = rng.random() < 0.3
    nz = int(rng.integers(1, 12 if small else max_size + 1))
    nx = int(rng.integers(1, 12 if small else max_size + 1))
    side = int(rng.choice([LEFT, RIGHT, TOP, BOTTOM]))
    contact = float(rng.choice([0.0, 0.0, 0.1, 0.3])) if min(nz, nx) >= 8 else 0.0
    kw = dict(outlet_side=side, roughness=float(rng.choice([0.0, 0.0, 0.25])),
              high_mask=float(rng.choice([100.0, 8.0])), avg_factor=float(rng.choice([1.0, 10.0])))
    mode = rng.choice(["pressure", "pressure", "back_pressure", "inlet_rate", "outlet_rate"])
    if mode == "back_pressure":
        kw.update(inlet_pressure=3.0e5 + 700.0, outlet_pressure=3.0e5)
    elif mode == "inlet_rate":
        kw.update(inlet_pressure=-1.0, outlet_pressure=50.0, inlet_rate=1e-9)
    elif mode == "outlet_rate":
        kw.update(inlet_pressure=500.0, outlet_pressure=-1.0, outlet_rate=1e-9)
    sigma_c = int(rng.choice([0, 2, 6])) if min(nz, nx) >= 4 else 0     # a smoothed 1 x 1 field has no variance to rescale
    amap = dict(nz=nz, nx=nx, seed=int(rng.integers(0, 1 << 30)), sigma_c=sigma_c,
                contact=contact, sigma_k=int(rng.choice([3, 6])))
    return amap, kw, mode


def has_open_path(a, low_mask, vertical):
    from scipy import ndimage
    lab, _ = ndimage.label(a > low_mask)
    first, last = (lab[0, :], lab[-1, :]) if vertical else (lab[:, 0], lab[:, -1])
    return bool((set(np.unique(first)) & set(np.unique(last))) - {0})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-size", type=int, default=160)
    ap.add_argument("--sim", action="store_true")
    args = ap.parse_args()
    import test_gpu_parity as tp
    if args.sim:                       # test infrastructure: the CPU emulation build in place of libapmlcl.so
        from sim import sim_lib
        from apmapflow_b200 import _lib
        _lib._lib = sim_lib.load()
    rng = np.random.default_rng(args.seed)
    t0, done, failed = time.time(), 0, 0
    while time.time() - t0 < args.seconds:
        amap, kw, mode = draw_case(rng, args.max_size)
        a = synth_aperture(amap["nz"], amap["nx"], seed=amap["seed"], sigma_c=amap["sigma_c"], contact=amap["contact"],
                           sigma_k=amap["sigma_k"])
        vertical = kw["outlet_side"] in (TOP, BOTTOM)
        pmask = anchored_open_mask(a, 1e-4, vertical) if amap["contact"] else None
        if amap["contact"] and not has_open_path(a, 1e-4, vertical):
            continue                   # blocked fracture: see the module docstring
        try:
            # back pressure: the stored pressures carry eps * |p|, the same floor the reference's FRAC_FLOW sees
            tol = 1e-6 if mode != "back_pressure" else 5e-6
            tp.check_against_oracle(tp.synth_params(**kw), a, ptol=tol * (5 if kw["roughness"] else 1), pmask=pmask)
        except Exception as exc:
            failed += 1
            last = traceback.extract_tb(exc.__traceback__)[-1]
            print("FAIL map=%s params=%s mode=%s: %s at %s:%d %s" %
                  (amap, kw, mode, type(exc).__name__, os.path.basename(last.filename), last.lineno,
                   str(exc).split("\n")[0][:200]), flush=True)
        done += 1
    print("fuzz parity: %d cases in %.1f s, %d failed (seed %d)" % (done, time.time() - t0, failed, args.seed), flush=True)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
