"""ctypes binding of libapmlcl.so (include/apm_lcl.h).

This is the stub the reference side would add in place of the ``Popen((exe, inp))`` in
``apmapflow/run_model/run_model.py:752-759``.  There is NO CPU fallback: if the CUDA library is
missing or fails to load, importing a compute entry point raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# APM_LCL_LIB: developer knob for A/B runs of two builds of the SAME library (never a different implementation)
LIB_PATH = os.environ.get("APM_LCL_LIB") or os.path.join(os.path.dirname(_HERE), "libapmlcl.so")

LEFT, RIGHT, TOP, BOTTOM = 0, 1, 2, 3
SIDE_NAMES = ("LEFT", "RIGHT", "TOP", "BOTTOM")


def side_from_string(text):
    """Substring match in the reference's order LEFT, RIGHT, TOP, BOTTOM (APM_MODULE.F:306-368)."""
    text = str(text).upper()
    for code, name in enumerate(SIDE_NAMES):
        if name in text:
            return code
    raise ValueError(" INVALID OUTLET SIDE ENTERED: " + text)


class Params(C.Structure):
    _fields_ = [
        ("inlet_pressure", C.c_double), ("outlet_pressure", C.c_double),
        ("inlet_rate", C.c_double), ("outlet_rate", C.c_double),
        ("density", C.c_double), ("viscosity", C.c_double),
        ("voxel_size", C.c_double), ("avg_factor", C.c_double),
        ("roughness", C.c_double), ("high_mask", C.c_double), ("low_mask", C.c_double),
        ("outlet_side", C.c_int),
        ("rel_tol", C.c_double), ("max_iter", C.c_int), ("flags", C.c_int),
    ]

    @classmethod
    def defaults(cls, **kw):
        """INITIALIZE_RUN defaults (APERTURE_MAP_FLOW.F:88-126)."""
        p = cls(inlet_pressure=-1.0, outlet_pressure=-1.0, inlet_rate=-1.0, outlet_rate=-1.0,
                density=1.0, viscosity=-1.0, voxel_size=1.0, avg_factor=1.0, roughness=0.0,
                high_mask=1e6, low_mask=1e-6, outlet_side=TOP, rel_tol=0.0, max_iter=0, flags=0)
        for k, v in kw.items():
            if k == "outlet_side" and isinstance(v, str):
                v = side_from_string(v)
            setattr(p, k, v)
        return p


_DP = C.POINTER(C.c_double)


class Fields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("pressure", "flow_x", "flow_z", "flow_m", "aperture", "reynolds", "residual")]


class Scalars(C.Structure):
    _fields_ = [
        ("zeros", C.c_longlong), ("masked", C.c_longlong),
        ("raw_aper", C.c_double), ("avg_aper", C.c_double),
        ("min_aper", C.c_double), ("max_aper", C.c_double),
        ("inlet_pressure", C.c_double), ("outlet_pressure", C.c_double),
        ("flow", C.c_double), ("outlet_rate", C.c_double), ("max_residual", C.c_double),
        ("iterations", C.c_int), ("converged", C.c_int),
        ("rel_residual", C.c_double), ("rel_residual_scaled", C.c_double),
        ("setup_ms", C.c_double), ("solve_ms", C.c_double), ("total_ms", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


EXPORTS = (
    "apm_lcl_version", "apm_lcl_last_error", "apm_lcl_last_launch_count", "apm_lcl_default_params",
    "apm_lcl_solver_create", "apm_lcl_solver_destroy", "apm_lcl_solver_solve", "apm_lcl_solve",
    "apm_lcl_conductance", "apm_lcl_solver_device_field", "apm_lcl_solver_time_kernel",
    "apm_lcl_run_file", "apm_lcl_bulk_solve",
    "apm_lcl_peer_handle_bytes", "apm_lcl_solver_create_striped", "apm_lcl_solver_connect",
    "apm_lcl_solver_disconnect", "apm_lcl_solver_rows", "apm_lcl_solve_striped",
    "apm_lcl_run_file_striped", "apm_lcl_percentiles", "apm_lcl_map_stats",
)

_lib = None


def load():
    """Load libapmlcl.so (once).  Raises OSError with a build hint when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(
            "libapmlcl.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    _lib = bind(lib)
    return _lib


def bind(lib):
    """Attach the prototypes of include/apm_lcl.h to a loaded library object."""
    lib.apm_lcl_version.restype = C.c_int
    lib.apm_lcl_last_error.restype = C.c_char_p
    lib.apm_lcl_last_launch_count.restype = C.c_longlong
    lib.apm_lcl_default_params.argtypes = [C.POINTER(Params)]
    lib.apm_lcl_default_params.restype = None
    lib.apm_lcl_solver_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
    lib.apm_lcl_solver_destroy.argtypes = [C.c_void_p]
    lib.apm_lcl_peer_handle_bytes.restype = C.c_int
    lib.apm_lcl_solver_create_striped.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.c_void_p]
    lib.apm_lcl_solver_connect.argtypes = [C.c_void_p, C.c_void_p]
    lib.apm_lcl_solver_disconnect.argtypes = [C.c_void_p]
    lib.apm_lcl_solver_rows.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
    lib.apm_lcl_solve_striped.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                          C.POINTER(Fields), C.POINTER(Scalars)]
    lib.apm_lcl_solver_solve.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p,
                                         C.POINTER(Fields), C.POINTER(Scalars)]
    lib.apm_lcl_solve.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(Fields), C.POINTER(Scalars)]
    lib.apm_lcl_conductance.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Scalars)]
    lib.apm_lcl_solver_device_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p),
                                                C.POINTER(C.c_size_t)]
    lib.apm_lcl_solver_time_kernel.argtypes = [C.c_void_p, C.c_char_p, C.c_int,
                                               C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.apm_lcl_bulk_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Params),
                                       C.POINTER(C.c_void_p), C.POINTER(Fields), C.POINTER(Scalars)]
    lib.apm_lcl_bulk_solve.restype = C.c_int
    lib.apm_lcl_run_file.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    lib.apm_lcl_run_file.restype = C.c_int
    lib.apm_lcl_run_file_striped.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_size_t]
    lib.apm_lcl_run_file_striped.restype = C.c_int
    lib.apm_lcl_percentiles.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.apm_lcl_map_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 5
    return lib


def last_error():
    return (load().apm_lcl_last_error() or b"").decode("utf-8", "replace")


class LclError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise LclError(last_error() or "libapmlcl call failed (rc=%d)" % rc)
