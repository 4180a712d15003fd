"""TEST INFRASTRUCTURE ONLY: loads the CPU emulation build of the kernel sources (see cuda_sim.h)."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_sim  # noqa: E402
from apmapflow_b200 import _lib  # noqa: E402

_sim = None


def load():
    global _sim
    if _sim is None:
        _sim = _lib.bind(C.CDLL(build_sim.build()))
    return _sim
