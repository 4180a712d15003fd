"""Tensor-level entry point of the B200 LCL solve (thin wrapper over the C ABI).

PyTorch appears here only as a container for device memory: a CUDA tensor contributes its
``data_ptr()``, results come back as CUDA tensors; numpy in -> numpy out (host buffers, the copies
happen inside the library).  The numerics live in libapmlcl.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Fields, Params, Scalars

FIELD_NAMES = ("pressure", "flow_x", "flow_z", "flow_m", "aperture", "reynolds", "residual")


def _is_torch_cuda(a):
    return type(a).__module__.startswith("torch") and getattr(a, "is_cuda", False)


class LclSolver(object):
    """Persistent solver for maps of one (nz, nx) on one device (apm_lcl_solver_create)."""

    def __init__(self, nx, nz, device=0, lib=None):
        self.lib = lib if lib is not None else _lib.load()
        self.nx, self.nz, self.device = int(nx), int(nz), int(device)
        self._h = C.c_void_p()
        rc = self.lib.apm_lcl_solver_create(C.byref(self._h), self.device, self.nx, self.nz)
        if rc != 0:
            raise _lib.LclError((self.lib.apm_lcl_last_error() or b"").decode())

    def close(self):
        if self._h:
            self.lib.apm_lcl_solver_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def solve(self, params, aperture, fields=FIELD_NAMES):
        """Runs the hot path.  Returns (dict name -> array [nz, nx], dict of scalars)."""
        out = Fields()
        arrays = {}
        if _is_torch_cuda(aperture):
            import torch
            a = aperture.to(torch.float64).contiguous()
            assert tuple(a.shape) == (self.nz, self.nx)
            ptr = a.data_ptr()
            for name in fields:
                arrays[name] = torch.empty_like(a)
                setattr(out, name, arrays[name].data_ptr())
            torch.cuda.current_stream(a.device).synchronize()   # inputs complete before our stream reads
        else:
            a = np.ascontiguousarray(aperture, dtype=np.float64)
            assert a.shape == (self.nz, self.nx)
            ptr = a.ctypes.data
            for name in fields:
                arrays[name] = np.empty((self.nz, self.nx))
                setattr(out, name, arrays[name].ctypes.data)
        sc = Scalars()
        rc = self.lib.apm_lcl_solver_solve(self._h, C.byref(params), ptr, C.byref(out), C.byref(sc))
        d = sc.as_dict()
        d["gpu_launches"] = int(self.lib.apm_lcl_last_launch_count())
        if rc != 0:
            err = _lib.LclError((self.lib.apm_lcl_last_error() or b"").decode())
            err.scalars = d
            err.fields = arrays
            raise err
        return arrays, d

    def time_kernel(self, name, reps=20):
        ms, nbytes = C.c_double(), C.c_double()
        rc = self.lib.apm_lcl_solver_time_kernel(self._h, name.encode(), int(reps), C.byref(ms), C.byref(nbytes))
        if rc != 0:
            raise _lib.LclError((self.lib.apm_lcl_last_error() or b"").decode())
        return ms.value, nbytes.value


def lcl_solve(params, aperture, device=0, fields=FIELD_NAMES, lib=None):
    """One-shot solve of a single map."""
    nz, nx = aperture.shape
    with LclSolver(nx, nz, device=device, lib=lib) as s:
        return s.solve(params, aperture, fields=fields)


def conductance(params, aperture, device=0, lib=None):
    """COEF + FLOW_BOUNDARY only: returns width[nz,nx], trx[nz,nx+1], trz[nz+1,nx], scalars (host arrays)."""
    lib = lib if lib is not None else _lib.load()
    a = np.ascontiguousarray(aperture, dtype=np.float64)
    nz, nx = a.shape
    width, trx, trz = np.empty((nz, nx)), np.empty((nz, nx + 1)), np.empty((nz + 1, nx))
    sc = Scalars()
    rc = lib.apm_lcl_conductance(C.byref(params), a.ctypes.data, nx, nz, device, width.ctypes.data,
                                 trx.ctypes.data, trz.ctypes.data, C.byref(sc))
    if rc != 0:
        raise _lib.LclError((lib.apm_lcl_last_error() or b"").decode())
    return width, trx, trz, sc.as_dict()


def bulk_solve(cases, devices=(0,), fields=("pressure",), lib=None):
    """Independent cases sharded over GPUs, one host thread per device (ctypes drops the GIL).

    ``cases`` is a list of ``(Params, aperture[nz, nx])`` with one common map shape (host numpy
    arrays).  Device ``d`` of ``devices`` gets the contiguous block ``dist.shard(cases, d, len(devices))``
    -- no exchange between cases, exactly like the reference's BulkRun.  Returns a list of
    ``(dict of arrays, dict of scalars)`` in case order."""
    import threading
    from . import dist
    lib = lib if lib is not None else _lib.load()
    if not cases:
        return []
    nz, nx = cases[0][1].shape
    results = [None] * len(cases)
    index = list(range(len(cases)))
    errors = []

    def work(slot, device):
        mine = dist.shard(index, slot, len(devices))
        if not mine:
            return
        n = len(mine)
        params = (Params * n)(*[cases[k][0] for k in mine])
        maps = [np.ascontiguousarray(cases[k][1], dtype=np.float64) for k in mine]
        assert all(m.shape == (nz, nx) for m in maps)
        ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in maps])
        outs = (Fields * n)()
        arrays = []
        for j in range(n):
            arr = {name: np.empty((nz, nx)) for name in fields}
            for name, a in arr.items():
                setattr(outs[j], name, a.ctypes.data)
            arrays.append(arr)
        scal = (Scalars * n)()
        failed = lib.apm_lcl_bulk_solve(int(device), nx, nz, n, params, ptrs, outs, scal)
        if failed:
            errors.append((device, failed, (lib.apm_lcl_last_error() or b"").decode()))
        for j, k in enumerate(mine):
            results[k] = (arrays[j], scal[j].as_dict())

    threads = [threading.Thread(target=work, args=(slot, dev)) for slot, dev in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise _lib.LclError("bulk_solve: %s" % errors)
    return results
