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


def _check_shape(shape, expected):
    """A wrong shape would make the library read past the buffer: refuse it here (not an assert, those vanish with -O)."""
    if tuple(shape) != tuple(expected):
        raise ValueError("aperture map has shape %s, the solver was created for %s" % (tuple(shape), tuple(expected)))


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
        if getattr(self, "_h", None):
            self.lib.apm_lcl_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the modules this needs may already be gone
            pass

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
            _check_shape(a.shape, (self.nz, self.nx))
            ptr = a.data_ptr()
            for name in fields:
                arrays[name] = torch.empty_like(a)
                setattr(out, name, arrays[name].data_ptr())
            torch.cuda.current_stream(a.device).synchronize()   # inputs complete before our stream reads
        else:
            a = np.ascontiguousarray(aperture, dtype=np.float64)
            _check_shape(a.shape, (self.nz, self.nx))
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


class StripedLclSolver(object):
    """One map cut into row strips over the ranks of a ``torch.distributed`` group, one process and one GPU
    per rank on ONE node (apm_lcl_solver_create_striped / _connect).  torch.distributed only
    carries the peer handles at start-up and the barrier at shut-down: ghost rows and dot products move
    between the GPUs inside the solver's kernels.  ``solve`` is collective."""

    def __init__(self, nx, nz, device=0, group=None, lib=None):
        import torch.distributed as dist
        self.lib = lib if lib is not None else _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nx, self.nz, self.device = int(nx), int(nz), int(device)
        nbytes = int(self.lib.apm_lcl_peer_handle_bytes())
        handle = C.create_string_buffer(nbytes)
        self._h = C.c_void_p()
        rc = self.lib.apm_lcl_solver_create_striped(C.byref(self._h), self.device, self.nx, self.nz, self.rank,
                                                    self.world, handle)
        # every rank reports, so that a failure on one does not leave the others waiting in the gather
        status = [None] * self.world
        dist.all_gather_object(status, (rc, handle.raw, self._error()), group=group)
        bad = [(k, st[2]) for k, st in enumerate(status) if st[0] != 0]
        if bad:
            self.close(barrier=False)
            raise _lib.LclError("striped solver: rank %d: %s" % bad[0])
        blob = b"".join(st[1] for st in status)
        rc = self.lib.apm_lcl_solver_connect(self._h, blob)
        status = [None] * self.world
        dist.all_gather_object(status, (rc, self._error()), group=group)
        bad = [(k, st[1]) for k, st in enumerate(status) if st[0] != 0]
        if bad:
            self.close(barrier=False)
            raise _lib.LclError("striped solver: rank %d could not map its peers: %s" % bad[0])
        z = [C.c_int() for _ in range(4)]
        self.lib.apm_lcl_solver_rows(self._h, *[C.byref(v) for v in z])
        self.z_lo, self.z_hi, self.a_lo, self.a_hi = [v.value for v in z]

    def _error(self):
        return (self.lib.apm_lcl_last_error() or b"").decode("utf-8", "replace")

    def close(self, barrier=True):
        """Collective: unmap the peers, wait for everyone, free the slab (a slab must outlive its mappings)."""
        if getattr(self, "_h", None):
            self.lib.apm_lcl_solver_disconnect(self._h)
            if barrier:
                import torch.distributed as dist
                dist.barrier(group=self.group)
            self.lib.apm_lcl_solver_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close(barrier=exc[0] is None)

    def solve(self, params, aperture, fields=FIELD_NAMES):
        """``aperture`` is the whole map [nz, nx] (the rank's rows are taken from it) or just the rows
        [a_lo, a_hi) this rank needs; numpy in -> numpy out, CUDA tensor in -> CUDA tensors out.  Returns
        (dict name -> array of the rows [z_lo, z_hi) this rank owns, dict of whole-map scalars)."""
        rows_in, rows_out = self.a_hi - self.a_lo, self.z_hi - self.z_lo
        if aperture.shape[0] == self.nz and rows_in != self.nz:
            aperture = aperture[self.a_lo:self.a_hi]
        _check_shape(aperture.shape, (rows_in, self.nx))
        out = Fields()
        arrays = {}
        if _is_torch_cuda(aperture):
            import torch
            a = aperture.to(torch.float64).contiguous()
            ptr = a.data_ptr()
            for name in fields:
                arrays[name] = torch.empty((rows_out, self.nx), dtype=torch.float64, device=a.device)
                setattr(out, name, arrays[name].data_ptr())
            torch.cuda.current_stream(a.device).synchronize()
        else:
            a = np.ascontiguousarray(aperture, dtype=np.float64)
            ptr = a.ctypes.data
            for name in fields:
                arrays[name] = np.empty((rows_out, self.nx))
                setattr(out, name, arrays[name].ctypes.data)
        sc = Scalars()
        rc = self.lib.apm_lcl_solver_solve(self._h, C.byref(params), ptr, C.byref(out), C.byref(sc))
        d = sc.as_dict()
        d["gpu_launches"] = int(self.lib.apm_lcl_last_launch_count())
        if rc != 0:
            err = _lib.LclError("rank %d: %s" % (self.rank, self._error()))
            err.scalars = d
            err.fields = arrays
            raise err
        return arrays, d

    def time_kernel(self, name, reps=20):
        """Collective: every rank times the kernel on its own strip (the dot products inside still meet)."""
        return LclSolver.time_kernel(self, name, reps)

    def gather(self, arrays):
        """Whole-map host arrays on every rank from the per-rank strips (convenience for tests and writers)."""
        import torch.distributed as dist
        host = {k: (v.cpu().numpy() if _is_torch_cuda(v) else v) for k, v in arrays.items()}
        parts = [None] * self.world
        dist.all_gather_object(parts, (self.z_lo, self.z_hi, host), group=self.group)
        if self.z_hi - self.z_lo == self.nz:          # nothing was striped: every rank holds the whole map
            return host
        return {k: np.concatenate([p[2][k] for p in sorted(parts, key=lambda p: p[0])], axis=0) for k in host}


class StripedGroup(object):
    """One map cut into row strips over several GPUs driven by ONE process: a solver per device, a host
    thread per solver (ctypes drops the GIL), the peers addressed directly instead of through CUDA IPC.
    ``devices`` may name a GPU more than once (strips then share it: a way to run the striped path on a
    single-GPU box)."""

    def __init__(self, nx, nz, devices, lib=None):
        self.lib = lib if lib is not None else _lib.load()
        self.nx, self.nz, self.devices = int(nx), int(nz), [int(d) for d in devices]
        world = len(self.devices)
        nbytes = int(self.lib.apm_lcl_peer_handle_bytes())
        self._h, blobs = [], []
        try:
            for rank, dev in enumerate(self.devices):
                h, blob = C.c_void_p(), C.create_string_buffer(nbytes)
                _lib_check(self.lib, self.lib.apm_lcl_solver_create_striped(C.byref(h), dev, self.nx, self.nz, rank, world, blob))
                self._h.append(h)
                blobs.append(blob.raw)
            for h in self._h:
                _lib_check(self.lib, self.lib.apm_lcl_solver_connect(h, b"".join(blobs)))
        except Exception:
            self.close()
            raise
        self.rows = []
        for h in self._h:
            z = [C.c_int() for _ in range(4)]
            self.lib.apm_lcl_solver_rows(h, *[C.byref(v) for v in z])
            self.rows.append(tuple(v.value for v in z))

    def close(self):
        for h in self._h:
            self.lib.apm_lcl_solver_disconnect(h)
        for h in self._h:
            self.lib.apm_lcl_solver_destroy(h)
        self._h = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def solve(self, params, aperture, fields=FIELD_NAMES):
        """``aperture``: whole map [nz, nx], host numpy array.  Returns (dict of whole-map arrays, scalars)."""
        import threading
        a = np.ascontiguousarray(aperture, dtype=np.float64)
        _check_shape(a.shape, (self.nz, self.nx))
        whole = {name: np.empty((self.nz, self.nx)) for name in fields}
        results = [None] * len(self._h)

        def work(k):
            z_lo, z_hi, a_lo, a_hi = self.rows[k]
            out = Fields()
            for name in fields:
                setattr(out, name, whole[name][z_lo:z_hi].ctypes.data)     # row slices of a C array are contiguous
            sc = Scalars()
            rc = self.lib.apm_lcl_solver_solve(self._h[k], C.byref(params), a[a_lo:a_hi].ctypes.data, C.byref(out), C.byref(sc))
            d = sc.as_dict()
            d["gpu_launches"] = int(self.lib.apm_lcl_last_launch_count())
            results[k] = (rc, d, (self.lib.apm_lcl_last_error() or b"").decode("utf-8", "replace"))

        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(self._h))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        bad = [(k, r[2]) for k, r in enumerate(results) if r[0] != 0]
        if bad:
            err = _lib.LclError("striped solve: rank %d: %s" % bad[0])
            err.scalars = [r[1] for r in results]
            err.fields = whole
            raise err
        d = dict(results[0][1])
        d["gpu_launches"] = sum(r[1]["gpu_launches"] for r in results)
        d["per_rank"] = [r[1] for r in results]
        return whole, d


def _lib_check(lib, rc):
    if rc != 0:
        raise _lib.LclError((lib.apm_lcl_last_error() or b"").decode("utf-8", "replace"))


def lcl_solve_striped(params, aperture, devices, fields=FIELD_NAMES, lib=None):
    """One-shot solve of one map striped over ``devices`` from a single process."""
    nz, nx = aperture.shape
    with StripedGroup(nx, nz, devices, lib=lib) as g:
        return g.solve(params, aperture, fields=fields)


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


def percentiles(values, pct, device=0, lib=None):
    """CALC_PERCENTILE (OUTPUT_MODULE.F:53-124) of a host array or CUDA tensor, on the device."""
    lib = lib if lib is not None else _lib.load()
    pct = np.ascontiguousarray(pct, dtype=np.float64)
    out = np.empty(len(pct))
    if _is_torch_cuda(values):
        v = values.contiguous()
        ptr, n = v.data_ptr(), v.numel()
    else:
        v = np.ascontiguousarray(values, dtype=np.float64)
        ptr, n = v.ctypes.data, v.size
    _lib_check(lib, lib.apm_lcl_percentiles(ptr, n, pct.ctypes.data_as(C.POINTER(C.c_double)), len(pct), int(device),
                                            out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def map_stats(data_map, device=0, lib=None):
    """MAP_STATS (MAP_MODULE.F:182-303): dict with avg, rms, dev and the x / z correlation lengths in cells."""
    lib = lib if lib is not None else _lib.load()
    if _is_torch_cuda(data_map):
        m = data_map.contiguous()
        ptr = m.data_ptr()
    else:
        m = np.ascontiguousarray(data_map, dtype=np.float64)
        ptr = m.ctypes.data
    nz, nx = m.shape
    vals = [C.c_double() for _ in range(5)]
    _lib_check(lib, lib.apm_lcl_map_stats(ptr, nx, nz, int(device), *[C.byref(v) for v in vals]))
    return dict(zip(("avg", "rms", "dev", "x_corr", "z_corr"), [v.value for v in vals]))


def bulk_solve(cases, devices=(0,), fields=("pressure",), lib=None):
    """Independent cases sharded over GPUs, one host thread per device (ctypes drops the GIL).

    ``cases`` is a list of ``(Params, aperture[nz, nx])`` with one common map shape (host numpy
    arrays).  Device ``d`` of ``devices`` gets the contiguous block ``dist.shard(cases, d, len(devices))``
    -- no exchange between cases, exactly like the reference's BulkRun.  Cases that pass the SAME array object and
    differ only in the two boundary pressures share one unit solution inside the library (the derived ones report
    ``iterations == 0``; ``APM_LCL_BULK_SHARE=0`` or ``params.flags = 1`` turns that off, see include/apm_lcl.h).
    Returns a list of ``(dict of arrays, dict of scalars)`` in case order."""
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
        try:
            run(slot, device)
        except Exception as exc:           # an exception on a worker thread must reach the caller
            errors.append((device, -1, "%s: %s" % (type(exc).__name__, exc)))

    def run(slot, device):
        mine = dist.shard(index, slot, len(devices))
        if not mine:
            return
        n = len(mine)
        params = (Params * n)(*[cases[k][0] for k in mine])
        maps = [np.ascontiguousarray(cases[k][1], dtype=np.float64) for k in mine]
        for m in maps:
            _check_shape(m.shape, (nz, nx))
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
