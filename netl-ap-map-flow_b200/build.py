"""Builds libapmlcl.so (and the apm-lcl-model.exe front end) with nvcc for sm_100a, in-tree.

    python netl-ap-map-flow_b200/build.py [--force] [--verbose]

The .so lands next to this file so it travels with the repository snapshot to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libapmlcl.so")
EXE = os.path.join(HERE, "apm-lcl-model.exe")
OBJ = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-extended-lambda", "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++"]


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") or (f.endswith(".cpp") and f != "main.cpp"))
    return [os.path.join(CSRC, f) for f in cu]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "apm_lcl.h"))
    return hs


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [NVCC] + ARCH + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-x", "cu", "-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-Xlinker", "-Bsymbolic", "-o", LIB] + objs + \
              ["-cudart", "static", "-ccbin", "/usr/bin/g++"]
        subprocess.check_call(cmd)
    main_cpp = os.path.join(CSRC, "main.cpp")
    if os.path.exists(main_cpp) and (force or _newer(EXE, [main_cpp, LIB])):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", main_cpp, "-o", EXE, "-L" + HERE, "-lapmlcl",
                               "-Wl,-rpath,$ORIGIN", "-I" + os.path.join(os.path.dirname(HERE), "include")])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
