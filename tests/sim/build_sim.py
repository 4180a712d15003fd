"""TEST INFRASTRUCTURE ONLY.  Compiles the kernel sources of libapmlcl for the CPU with the CUDA
emulation shim (cuda_sim.h) so their logic can be checked against the oracle in a container
without a GPU.  Output: tests/sim/libapmlcl_sim.so -- loaded by tests only, never by the product.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "netl-ap-map-flow_b200", "csrc")
LIB = os.path.join(HERE, "libapmlcl_sim.so")
OBJ = os.path.join(HERE, "build")


def build(force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                  if f.endswith(".cu") or (f.endswith(".cpp") and f != "main.cpp"))
    srcs.append(os.path.join(HERE, "sim_globals.cpp"))
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))] + \
        [os.path.join(HERE, "cuda_sim.h"), os.path.join(ROOT, "include", "apm_lcl.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-DLCL_CPU_SIM=1",
               "-include", os.path.join(HERE, "cuda_sim.h"), "-I", HERE, "-x", "c++", "-c", s, "-o", o,
               "-ffp-contract=off", "-Wno-unused-function", "-Wno-unused-variable"]
        procs.append(subprocess.Popen(cmd))
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("sim build failed")
    subprocess.check_call(["/usr/bin/g++", "-shared", "-Wl,-Bsymbolic", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
