"""CPU-side checks of the drop-in boundary: libapmlcl.so loads and exports every symbol that
include/apm_lcl.h declares; nothing here launches a kernel."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "apm_lcl.h")


@pytest.fixture(scope="module")
def lib():
    from apmapflow_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "netl-ap-map-flow_b200", "build.py")])
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"APM_LCL_API[^;(]*?\b(apm_lcl_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("apm_lcl_solve", "apm_lcl_solver_create", "apm_lcl_solver_solve", "apm_lcl_solver_destroy",
                 "apm_lcl_conductance", "apm_lcl_last_error", "apm_lcl_version"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), "libapmlcl.so does not export %s" % name


def test_version_and_defaults(lib):
    from apmapflow_b200._lib import Params
    assert lib.apm_lcl_version() == 100
    p = Params()
    lib.apm_lcl_default_params(C.byref(p))
    d = Params.defaults()
    for name, _ in Params._fields_:
        assert getattr(p, name) == getattr(d, name), name
    # INITIALIZE_RUN defaults, APERTURE_MAP_FLOW.F:94-111
    assert (p.inlet_pressure, p.outlet_pressure, p.viscosity, p.high_mask, p.low_mask) == (-1.0, -1.0, -1.0, 1e6, 1e-6)


def test_struct_sizes_match_the_c_header(tmp_path):
    """sizeof() of the ctypes mirrors equals the C compiler's view of include/apm_lcl.h."""
    from apmapflow_b200._lib import Fields, Params, Scalars
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu\\n", sizeof(apm_lcl_params),'
                   ' sizeof(apm_lcl_fields), sizeof(apm_lcl_scalars));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(Params), C.sizeof(Fields), C.sizeof(Scalars)]


def test_product_fails_loudly_without_the_cuda_library(monkeypatch):
    from apmapflow_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libapmlcl.so")
    with pytest.raises(OSError):
        _lib.load()


def test_no_gpu_means_error_not_fallback(lib):
    """On a box without a CUDA device the compute entry points return 1 with a message."""
    import numpy as np
    from apmapflow_b200._lib import Params, Scalars
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    a = np.ones((4, 4))
    sc = Scalars()
    p = Params.defaults(inlet_pressure=1.0, outlet_pressure=0.0, viscosity=1.0)
    rc = lib.apm_lcl_solve(C.byref(p), a.ctypes.data, 4, 4, 0, None, C.byref(sc))
    assert rc == 1
    assert b"CUDA" in lib.apm_lcl_last_error()


def test_product_package_never_references_the_oracle_or_the_sim():
    pkg = os.path.join(ROOT, "netl-ap-map-flow_b200")
    for base, _, files in os.walk(pkg):
        if os.path.basename(base) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle_binding" not in text and "libapm_oracle" not in text, f
                assert "libapmlcl_sim" not in text and "sim_lib" not in text, f
