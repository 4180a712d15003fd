"""ctypes binding of the CPU oracle (oracle/libapm_oracle.so).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never from the
product package."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "netl-ap-map-flow_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from apmapflow_b200._lib import Params, Scalars  # noqa: E402  (struct layouts are shared)

LIB_PATH = os.path.join(_HERE, "libapm_oracle.so")


class OracleOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("width", "trx", "trz", "point_map", "pressure", "flow_x", "flow_z", "flow_m",
                 "reynolds", "residual", "q")]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "apm_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "CC=gcc"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        for name in ("apm_oracle_solve", "apm_oracle_solve_quad"):
            f = getattr(lib, name)
            f.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_int, C.POINTER(OracleOut),
                          C.POINTER(Scalars), C.POINTER(C.c_double)]
            f.restype = C.c_int
        lib.apm_oracle_map_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.apm_oracle_map_stats.restype = None
        lib.apm_oracle_d4_fill_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.apm_oracle_d4_fill_bytes.restype = C.c_double
        _lib = lib
    return _lib


def solve(params, aperture, quad=False, want=("width", "trx", "trz", "pressure", "flow_x", "flow_z",
                                             "flow_m", "reynolds", "residual")):
    """Run the oracle.  Returns (dict of numpy arrays, scalars dict incl. 'solve_seconds')."""
    lib = load()
    a = np.ascontiguousarray(aperture, dtype=np.float64)
    nz, nx = a.shape
    shapes = {"width": (nz, nx), "trx": (nz, nx + 1), "trz": (nz + 1, nx), "point_map": (nz, nx, 4),
              "pressure": (nz, nx), "flow_x": (nz, nx), "flow_z": (nz, nx), "flow_m": (nz, nx),
              "reynolds": (nz, nx), "residual": (nz, nx), "q": (5, nz * nx)}
    arrs = {k: np.zeros(shapes[k]) for k in want}
    out = OracleOut()
    for k, v in arrs.items():
        setattr(out, k, v.ctypes.data)
    sc = Scalars()
    secs = C.c_double(0.0)
    fn = lib.apm_oracle_solve_quad if quad else lib.apm_oracle_solve
    rc = fn(C.byref(params), a.ctypes.data, nx, nz, C.byref(out), C.byref(sc), C.byref(secs))
    if rc != 0:
        raise ValueError("oracle rejected the input (rc=%d)" % rc)
    d = sc.as_dict()
    d["solve_seconds"] = secs.value
    return arrs, d


def map_stats(a):
    lib = load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.zeros(5)
    lib.apm_oracle_map_stats(a.ctypes.data, a.shape[1], a.shape[0], out.ctypes.data)
    return dict(zip(("rms", "stddev", "x_corr", "z_corr", "mean"), out))
