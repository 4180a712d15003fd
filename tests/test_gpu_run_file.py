"""-m gpu: the file-level drop-in through libapmlcl.so and the Python run API
(reference tests: test/unit/TestRunCore.py:140-162, test/integration/TestScripts.py:45-83,
test/fortran/bin/test_model_code_coverage)."""
import os
import subprocess
import time

import pytest

from apmapflow_b200 import _lib
from apmapflow_b200.run_model import BulkRun, InputFile, run_model
from test_run_file_sim import GOLDEN, check_against_golden_files, coverage_inp, run

pytestmark = pytest.mark.gpu


def test_coverage_case_reproduces_golden_files_on_gpu(tmp_path):
    rc, out = run(_lib.load(), str(tmp_path), coverage_inp(), "coverage_test.inp")
    assert rc == 0, out
    check_against_golden_files(str(tmp_path))


def test_executable_front_end(tmp_path):
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "apm-lcl-model.exe")
    assert os.path.exists(exe)
    (tmp_path / "coverage_test.inp").write_text(coverage_inp())
    proc = subprocess.run([exe, "coverage_test.inp"], cwd=str(tmp_path), stdout=subprocess.PIPE, universal_newlines=True)
    assert proc.returncode == 0, proc.stdout
    check_against_golden_files(str(tmp_path))
    proc = subprocess.run([exe], cwd=str(tmp_path), stdout=subprocess.PIPE, universal_newlines=True)
    assert proc.returncode == 1 and "FATAL ERROR - NO INPUT FILE" in proc.stdout


def _model_inputs(tmp_path, amap="parallel-plate-01vox.txt"):
    inp = InputFile(os.path.join(GOLDEN, "test-model-inputs.txt"))
    for key in ("APER-MAP", "SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE"):
        inp[key].commented_out = False
    for key in ("SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE"):
        inp.filename_formats[key] = os.path.join(str(tmp_path), "extra", key + ".RCTEST")
    inp.filename_formats["input_file"] = os.path.join(str(tmp_path), "test-model-inputs.txt")
    inp["APER-MAP"] = os.path.join(GOLDEN, amap)
    return inp


def test_run_model_async_then_sync(tmp_path):           # TestRunCore.py:140-162
    inp = _model_inputs(tmp_path, "Fracture1ApertureMap-10avg.txt")
    proc = run_model(inp, synchronous=False)
    assert proc.poll() is None or proc.poll() == 0      # a 138x136 solve may already be done
    assert proc.wait() == 0
    assert proc.input_file is inp and proc.end_time >= proc.start_time
    assert "EXITING PROGRAM" in proc.stdout_content
    proc = run_model(inp, synchronous=True, show_stdout=True)
    assert proc.poll() == 0
    assert os.path.isfile(os.path.join(str(tmp_path), "extra", "SUMMARY-FILE.RCTEST"))
    assert os.path.isfile(os.path.join(str(tmp_path), "extra", "FLOW-FILE-x.RCTEST"))


def test_bulk_run_writes_every_case(tmp_path):          # TestScripts.py:45-83
    import yaml
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "test-bulk-run.yaml")))
    inp = InputFile(os.path.join(GOLDEN, cfg["initial_input_file"]))
    for key in ("APER-MAP", "SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE"):
        inp[key].commented_out = False
    fmts = {k: os.path.join(str(tmp_path), v) for k, v in cfg["default_file_formats"].items()}
    fmts["APER-MAP"] = os.path.join(GOLDEN, "{aper-map}.txt")
    kw = dict(cfg["bulk_run_keyword_args"], spawn_delay=0.0, retest_delay=0.01, sys_RAM=8.0, num_CPUs=4)
    bulk = BulkRun(inp, **kw)
    bulk.generate_input_files(cfg["default_run_parameters"], fmts, cfg["case_identifier"], cfg["case_parameters"])
    assert len(bulk.input_file_list) == 16
    t0 = time.time()
    bulk.start()
    bulk.wait()
    assert time.time() - t0 < 120
    logs = [f for f in os.listdir(str(tmp_path)) if f.endswith("-LOG.TXT")]
    vtks = [f for f in os.listdir(str(tmp_path)) if f.endswith("-VTK.vtk")]
    stats = [f for f in os.listdir(str(tmp_path)) if f.endswith("-STAT.yaml")]
    assert len(logs) == len(vtks) == len(stats) == 16
    st = yaml.safe_load(open(os.path.join(str(tmp_path), "parallel-plate-10vox-RF0.00-800PA-STAT.yaml")))
    assert st["OUTLET PRESS"] == [800.0, "PA"] and st["OUTLET RATE"][1] == "MM^3/MIN"
