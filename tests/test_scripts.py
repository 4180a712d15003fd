"""Command-line callers of the path (apmapflow_b200.scripts.apm_run_lcl_model / apm_bulk_run), after the
reference's test/integration/TestScripts.py:45-83 (bulk run, dry run then --start) and :157-186 (single runs).
The CPU variants drive the host logic through the CPU emulation build of the kernels (tests/sim, test
infrastructure only); the -m gpu variants go through libapmlcl.so."""
import os
import re
import sys

import pytest
import yaml

from apmapflow_b200 import _lib
from apmapflow_b200.run_model import InputFile
from apmapflow_b200.scripts import apm_bulk_run, apm_run_lcl_model

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OUTPUT_KEYS = ("SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE")


@pytest.fixture
def sim_library(monkeypatch):
    """Points the package's library handle at the CPU emulation build for the duration of a test."""
    from sim import sim_lib
    monkeypatch.setattr(_lib, "_lib", sim_lib.load())


def _single_run_input(tmp_path):
    inp = InputFile(os.path.join(GOLDEN, "test-model-inputs.txt"),
                    {"input_file": os.path.join(str(tmp_path), "model-input-file.inp")})
    inp["APER-MAP"] = os.path.join(GOLDEN, "Fracture1ApertureMap-10avg.txt")
    files = {}
    for key, arg in inp.items():
        if re.search(r"FILE$", key):
            files[key] = os.path.join(str(tmp_path), arg.value)
            inp[key] = files[key]
    for key in ("APER-MAP",) + OUTPUT_KEYS:
        inp[key].commented_out = False
    inp.write_inp_file()
    del files["FLOW-FILE"], files["STAT-FILE"]      # written as -x/-z/-m and .csv/.yaml, not under the plain name
    return inp, files


def _check_single_run(tmp_path, capsys):
    inp, files = _single_run_input(tmp_path)
    assert apm_run_lcl_model.main(["-v", inp.outfile_name]) == 0
    for path in files.values():
        assert os.path.isfile(path), path
    root, ext = os.path.splitext(os.path.join(str(tmp_path), inp["FLOW-FILE"].value))
    assert all(os.path.isfile(root + tag + ext) for tag in ("-x", "-z", "-m"))
    assert "EXITING PROGRAM" in capsys.readouterr().out          # show_stdout=True like the reference
    # a second run without OVERWRITE must fail like the model does, and the script reports it
    inp["OVERWRITE"].commented_out = True
    inp.write_inp_file()
    assert apm_run_lcl_model.main([inp.outfile_name]) == 1
    assert re.search(r'SPECIFY "OVERWRITE" IN THE', capsys.readouterr().out)


def _bulk_yaml(tmp_path, concurrent, trimmed=False):
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "test-bulk-run.yaml")))
    if trimmed:       # two cases instead of sixteen: a run through the CPU emulation takes ~15 s
        cfg["default_run_parameters"].update({"aper-map": ["parallel-plate-10vox"], "ROUGHNESS": ["0.00"]})
        cfg["case_parameters"] = {"parallel-plate-10vox": {"OUTLET-PRESS": ["800", "700"]}}
    init = InputFile(os.path.join(GOLDEN, cfg["initial_input_file"]))
    for key in ("APER-MAP",) + OUTPUT_KEYS:
        init[key].commented_out = False
    init.filename_formats["input_file"] = os.path.join(str(tmp_path), "bulk-init.inp")
    init.write_inp_file()
    cfg["initial_input_file"] = init.outfile_name
    for key, fmt in cfg["default_file_formats"].items():
        base = GOLDEN if key == "APER-MAP" else str(tmp_path)
        cfg["default_file_formats"][key] = os.path.join(base, fmt)
    cfg["bulk_run_keyword_args"].update(spawn_delay=0.0, retest_delay=0.01, num_CPUs=concurrent)
    path = os.path.join(str(tmp_path), "test-bulk-run-trimmed.yaml" if trimmed else "test-bulk-run.yaml")
    with open(path, "w") as handle:
        handle.write(yaml.dump(cfg))
    return path


def _check_bulk_run(tmp_path, capsys, concurrent, trimmed=False):
    path = _bulk_yaml(tmp_path, concurrent)
    assert apm_bulk_run.main(["-v", path]) == 0                    # dry run: input files only
    assert 'Add "--start" flag' in capsys.readouterr().out
    names = os.listdir(str(tmp_path))
    assert sum(n.endswith("-INIT.INP") for n in names) == 16 and not any(n.endswith("-LOG.TXT") for n in names)
    if trimmed:
        path = _bulk_yaml(tmp_path, concurrent, trimmed=True)
    assert apm_bulk_run.main(["-v", "--start", "--wait", "--devices", "0", path]) == 0
    ran = 2 if trimmed else 16
    names = os.listdir(str(tmp_path))
    assert sum(n.endswith("-LOG.TXT") for n in names) == sum(n.endswith("-VTK.vtk") for n in names) == ran
    stat = yaml.safe_load(open(os.path.join(str(tmp_path), "parallel-plate-10vox-RF0.00-800PA-STAT.yaml")))
    assert stat["OUTLET PRESS"] == [800.0, "PA"]
    if trimmed:
        return
    for name in ("parallel-plate-01vox-RF1.00-400PA-STAT.csv", "parallel-plate-01vox-RF1.00-400PA-STAT.yaml",
                 "parallel-plate-10vox-RF1.00-700PA-LOG.TXT", "Fracture1ApertureMap-10avg-RF0.00-300PA-INIT.INP"):
        assert os.path.isfile(os.path.join(str(tmp_path), name)), name      # TestScripts.py:76-83


def test_run_lcl_model_script_sim(tmp_path, capsys, sim_library):
    _check_single_run(tmp_path, capsys)


def test_bulk_run_script_sim(tmp_path, capsys, sim_library):
    _check_bulk_run(tmp_path, capsys, concurrent=1, trimmed=True)      # the emulation keeps its fibers in globals: one run at a time


@pytest.mark.gpu
def test_run_lcl_model_script_gpu(tmp_path, capsys):
    _check_single_run(tmp_path, capsys)


@pytest.mark.gpu
def test_bulk_run_script_gpu(tmp_path, capsys):
    _check_bulk_run(tmp_path, capsys, concurrent=4)


def test_script_command_lines_match_the_reference():
    """Same options as apmapflow/scripts/apm_run_lcl_model.py:29-38 and apm_bulk_run.py:32-40, plus the device ones."""
    opts = {a.dest for a in apm_run_lcl_model.build_parser()._actions}
    assert {"verbose", "executable", "input_files", "device", "devices"} <= opts
    opts = {a.dest for a in apm_bulk_run.build_parser()._actions}
    assert {"verbose", "start", "input_files", "devices"} <= opts
    args = apm_run_lcl_model.build_parser().parse_args(["--devices", "0,1", "a.inp", "b.inp"])
    assert args.devices == [0, 1] and [os.path.basename(p) for p in args.input_files] == ["a.inp", "b.inp"]
