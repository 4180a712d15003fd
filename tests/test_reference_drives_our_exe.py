"""INTEGRATION.md way 1, checked end to end: the UNMODIFIED reference package (imported from /root/reference, so this
only runs in the authoring container) drives our `apm-lcl-model.exe` front end through its own `run_model` and
`BulkRun` -- Popen, `;EXE-FILE:` pseudo-comment, AsyncCommunicate, exit codes (run_model.py:453-491, 749-772;
bulk_run.py:107-128; after test/unit/TestRunCore.py:140-162).  The executable is csrc/main.cpp linked against the CPU
emulation build of the kernels (test infrastructure) because there is no GPU here; the product executable is the same
main.cpp linked against libapmlcl.so (tests/test_gpu_run_file.py::test_executable_front_end runs that one)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/apmapflow"), reason="reference tree not present")


@pytest.fixture(scope="module")
def sim_exe(tmp_path_factory):
    from sim import build_sim
    lib = build_sim.build()
    exe = str(tmp_path_factory.mktemp("simexe") / "apm-lcl-model.exe")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", os.path.join(ROOT, "netl-ap-map-flow_b200", "csrc", "main.cpp"),
                           "-o", exe, "-I" + os.path.join(ROOT, "include"), lib, "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, "/root/reference")
    try:
        import apmapflow.run_model as ref_rm
    except Exception as exc:
        pytest.skip("reference package not importable: %s" % exc)
    finally:
        sys.path.remove("/root/reference")
    return ref_rm


def _plate(path, value, nz=20, nx=16):
    with open(path, "w") as f:
        f.write("\n".join(" ".join(["%.1f" % value] * nx) for _ in range(nz)) + "\n")
    return path


def _input(ref, tmp_path, exe):
    inp = ref.InputFile(os.path.join(GOLDEN, "test-model-inputs.txt"))
    for key in ("APER-MAP", "SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE"):
        inp[key].commented_out = False
        if key != "APER-MAP":
            inp.filename_formats[key] = os.path.join(str(tmp_path), "out", key + ".RCTEST")
    inp.filename_formats["input_file"] = os.path.join(str(tmp_path), "model-inputs.inp")
    inp["APER-MAP"] = _plate(os.path.join(str(tmp_path), "plate.txt"), 2.0)
    inp.add_parameter(";EXE-FILE: " + exe)              # the reference's own way to name another executable
    assert inp.executable == exe
    return inp


def test_reference_run_model_drives_our_executable(ref, sim_exe, tmp_path):
    inp = _input(ref, tmp_path, sim_exe)
    proc = ref.run_model(inp, synchronous=False)
    assert proc.poll() is None or proc.poll() == 0
    assert proc.wait() == 0
    proc = ref.run_model(inp, synchronous=True)
    assert proc.poll() == 0 and proc.returncode == 0
    assert "**** EXITING PROGRAM ****" in proc.stdout_content and proc.end_time >= proc.start_time
    out = os.path.join(str(tmp_path), "out")
    for name in ("SUMMARY-FILE.RCTEST", "STAT-FILE.csv", "STAT-FILE.yaml", "APER-FILE.RCTEST", "FLOW-FILE-x.RCTEST",
                 "FLOW-FILE-z.RCTEST", "FLOW-FILE-m.RCTEST", "PRESS-FILE.RCTEST", "VTK-FILE.rctest"):
        assert os.path.isfile(os.path.join(out, name)), name
    # failure travels as exit status 1 + the reference's message on stdout (test_model_code_coverage:226-231)
    inp["OVERWRITE"].commented_out = True
    proc = ref.run_model(inp, synchronous=True)
    assert proc.returncode == 1 and 'SPECIFY "OVERWRITE" IN THE' in proc.stdout_content


def test_reference_bulk_run_drives_our_executable(ref, sim_exe, tmp_path):
    inp = _input(ref, tmp_path, sim_exe)
    formats = {k: os.path.join(str(tmp_path), "bulk", "{OUTLET-PRESS}PA-" + k) for k in
               ("SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE")}
    formats["input_file"] = os.path.join(str(tmp_path), "bulk", "{OUTLET-PRESS}PA-INIT.INP")
    bulk = ref.BulkRun(inp, num_CPUs=2, sys_RAM=1.0, spawn_delay=0.05, retest_delay=0.05)
    bulk.generate_input_files({"OUTLET-PRESS": ["100", "200", "300"], "INLET-PRESS": ["1000"]}, formats)
    assert len(bulk.input_file_list) == 3
    bulk.start()                       # returns when the queue is empty; the children may still be running
    import time
    import yaml
    deadline = time.time() + 120
    want = [os.path.join(str(tmp_path), "bulk", "%sPA-STAT-FILE.yaml" % p) for p in ("100", "200", "300")]
    logs = [w.replace("STAT-FILE.yaml", "SUMMARY-FILE") for w in want]

    def finished(path):                # the log file ends with this line (tests/golden/coverage_test_log-orig.txt), written once every output file is closed
        return os.path.isfile(path) and "OUTPUT FILES CLOSED" in open(path).read()
    while time.time() < deadline and not all(finished(p) for p in logs):
        time.sleep(0.2)
    rates = []
    for w in want:
        st = yaml.safe_load(open(w))
        rates.append(st["OUTLET RATE"][0])
        assert st["INLET PRESS"][0] == pytest.approx(1000.0)
    # parallel plate: the flow rate is proportional to the pressure drop
    assert rates[0] / rates[2] == pytest.approx(900.0 / 700.0, rel=1e-6)
    assert rates[1] / rates[2] == pytest.approx(800.0 / 700.0, rel=1e-6)
