"""Randomised check of the STRIPED solve on the CPU emulation (tests/sim; test infrastructure): random map shapes, 2-4
gloo ranks, outlet sides, boundary-condition modes, W-levels and striping thresholds, each compared with the single-rank
solve of the same map through the assertions of tests/test_striped_sim.py::compare.  One line per failure with the
parameters that reproduce it.  Blocked fractures (no open path from inlet to outlet) are skipped for the reason given in
tools/gpu_fuzz_parity.py: two iterative solves to 1e-10 agree on nothing but noise at the scale of the LOW-MASK leak.

    python tools/sim_fuzz_striped.py [--seconds 300] [--seed 0]
"""
import argparse
import os
import sys
import tempfile
import time
import traceback
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
os.environ["PYTHONPATH"] = os.pathsep.join(sys.path[:4] + [os.environ.get("PYTHONPATH", "")])   # spawned ranks
import numpy as np  # noqa: E402
from apmapflow_b200._lib import LEFT, RIGHT, TOP, BOTTOM  # noqa: E402
from gpu_fuzz_parity import has_open_path  # noqa: E402
from helpers import synth_aperture  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300.0)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import test_striped_sim as ts
    rng = np.random.default_rng(args.seed)
    t0, done, failed = time.time(), 0, 0
    while time.time() - t0 < args.seconds:
        world = int(rng.choice([2, 2, 3, 4]))
        nz = int(rng.integers(40, 400))
        nx = int(rng.integers(3, 120))
        side = int(rng.choice([LEFT, RIGHT, TOP, BOTTOM]))
        pkw = ts.params(outlet_side=side)
        mode = str(rng.choice(["pressure", "pressure", "back_pressure", "inlet_rate", "outlet_rate"]))
        if mode == "back_pressure":
            pkw.update(inlet_pressure=2.0e5 + 500.0, outlet_pressure=2.0e5)
        elif mode == "inlet_rate":
            pkw.update(inlet_pressure=-1.0, outlet_pressure=50.0, inlet_rate=1e-9)
        elif mode == "outlet_rate":
            pkw.update(inlet_pressure=500.0, outlet_pressure=-1.0, outlet_rate=1e-9)
        min_cells = int(rng.choice([200, 500, 2000, 1 << 20]))
        seed = int(rng.integers(0, 1 << 20))
        env_w = {}
        if rng.random() < 0.3:
            env_w = {"APM_LCL_W_LO": "1", "APM_LCL_W_HI": "2"}
        a = synth_aperture(nz, nx, seed=seed, sigma_c=4, contact=0.15, sigma_k=6)     # the map test_striped_sim's ranks build
        if not has_open_path(a, 1e-4, side in (TOP, BOTTOM)):
            continue
        case = dict(world=world, shape=(nz, nx), seed=seed, mode=mode, side=side, min_cells=min_cells, w=env_w)
        old = {k: os.environ.get(k) for k in env_w}
        os.environ.update(env_w)
        try:
            with tempfile.TemporaryDirectory() as tmp:
                whole, allsc = ts.run_striped(Path(tmp), world, (nz, nx), seed, pkw, min_cells)
            f1, s1 = ts.single((nz, nx), seed, pkw)
            ts.compare(whole, allsc, f1, s1, world)
        except BaseException as exc:
            if isinstance(exc, KeyboardInterrupt):
                raise
            failed += 1
            last = traceback.extract_tb(exc.__traceback__)[-1]
            print("FAIL %s: %s at %s:%d %s" % (case, type(exc).__name__, os.path.basename(last.filename), last.lineno,
                                               str(exc).split("\n")[0][:300]), flush=True)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        done += 1
    print("striped sim fuzz: %d cases in %.0f s, %d failed (seed %d)" % (done, time.time() - t0, failed, args.seed), flush=True)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
