"""Host-side mirror of the reference's run API (apmapflow_b200.run_model).  The cases follow the
reference's own unit tests (test/unit/TestRunCore.py, test/unit/TestBulkRun.py) so a user of the
reference finds the same behaviour; where /root/reference is present the two InputFile
implementations are also diffed live."""
import os
import sys

import pytest

from apmapflow_b200 import run_model as rm
from apmapflow_b200.run_model import ArgInput, BulkRun, InputFile

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_arg_input_parsing():          # TestRunCore.py:23-72
    arg = ArgInput(";INLET-PRESS:   100  KPA ;CMT-MSG")
    assert (arg.keyword, arg.value, arg.unit, arg.comment_msg, arg.commented_out) == \
        ("INLET-PRESS", "100", "KPA", ";CMT-MSG", True)
    assert arg.line == str(arg)
    arg.value = 200
    arg.unit = "PSI"
    assert (arg.value, arg.unit, arg.commented_out) == ("200", "PSI", False)
    arg = ArgInput(";OVERWRITE EXISTING FILES")
    assert arg.value == "OVERWRITE EXISTING FILES" and arg.unit is None
    arg.value = ("", True)
    assert arg.line == ";"
    arg = ArgInput("INLET-PRESS:")
    assert (arg.value, arg.unit, arg.commented_out) == ("", "", False)
    arg = ArgInput("")
    assert arg.value == arg.line
    assert ArgInput('KEYWORD: "VALUE1 + VALUE2"').value == '"VALUE1 + VALUE2"'
    arg = ArgInput('TEST: "UNMATCHED QUOTE')
    assert (arg.value, arg.unit) == ('"UNMATCHED', "QUOTE")


def test_input_file_clone_update_formats(tmp_path):      # TestRunCore.py:74-129
    inp = InputFile(os.path.join(GOLDEN, "test-model-inputs.txt"))
    assert inp.keys()
    clone = inp.clone()
    assert [inp[k].value for k in inp if k] == [clone[k].value for k in clone if k]
    InputFile(inp)
    inp.update({"INLET-PRESS": "300", "BAD-ARG": "IN FORMATS"})
    assert inp["INLET-PRESS"].value == "300"
    assert inp.filename_format_args["BAD-ARG"] == "IN FORMATS"
    with pytest.raises(TypeError):
        inp.update("a", "b", "c")
    exe = os.path.join(rm.DEFAULT_MODEL_PATH, rm.DEFAULT_MODEL_NAME)
    inp.add_parameter(";EXE-FILE: " + exe)
    assert inp["EXE-FILE"].value == exe
    assert [k for k, v in inp.items() if not v.commented_out] == list(inp.get_uncommented_values().keys())
    with pytest.raises(KeyError):
        inp.filename_formats["NONEXISTANT-FILE"] = "NONEXISTANT-FILE-FORMAT"
        print(inp)
    del inp.filename_formats["NONEXISTANT-FILE"]
    inp.filename_formats["input_file"] = "OUT.INP"
    inp.write_inp_file(alt_path=str(tmp_path))
    again = InputFile(str(tmp_path / "OUT.INP"))
    assert again["INLET-PRESS"].value == "300"
    # a non-existent EXE-FILE falls back to the packaged model (reference :453-491)
    again["EXE-FILE"] = "does-not-exist.exe"
    assert again.executable == exe


def test_estimate_req_ram(tmp_path):    # TestRunCore.py:131-138
    path = os.path.join(GOLDEN, "parallel-plate-01vox.txt")
    assert rm.map_dimensions(path) == (100, 100)
    assert rm.map_dimensions(os.path.join(GOLDEN, "Fracture1ApertureMap-10avg.txt")) == (136, 138)
    odd = tmp_path / "odd.txt"          # comment-only and blank records do not count, the last record gives the width
    odd.write_text("; header comment\n\n1, 2, 3 ; trailing\n4\t5\t6\n   \n7=8=9 10\n;end\n")
    assert rm.map_dimensions(str(odd)) == (4, 3)
    ram = rm.estimate_req_RAM([path], 10)
    assert 0 < ram[0] < 10
    with pytest.raises(EnvironmentError):
        rm.estimate_req_RAM([path], 0)
    assert rm.estimate_req_RAM([path], 0, suppress=True)


def test_bulk_run_defaults_and_combinations():          # TestBulkRun.py:21-44
    inp = InputFile(os.path.join(GOLDEN, "test-model-inputs.txt"))
    bulk = BulkRun(inp)
    assert not bulk.input_file_list
    assert (bulk.num_CPUs, bulk.sys_RAM, bulk.avail_RAM) == (2, 4.0, 3.6)
    combs = BulkRun._combine_run_params({"param1": [1, 2, 3], "param2": [4, 5], "param3": [6], "param4": None})
    assert len(combs) == 6
    assert set(combs[0].keys()) == {"param1", "param2", "param3"}
    assert set(combs[0].values()) == {6, 4, 1} and set(combs[-1].values()) == {6, 5, 3}


def test_bulk_run_check_processes():                    # TestBulkRun.py:46-67
    class Proc:
        value = None

        def poll(self):
            if self.value is None:
                self.value = 0
                return None
            return self.value
    procs, ram = [Proc()], [0.0]
    BulkRun._check_processes(procs, ram, retest_delay=0)
    assert not procs and not ram


def test_bulk_run_generate_input_files():               # TestBulkRun.py:69-94 + test-bulk-run.yaml
    import yaml
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "test-bulk-run.yaml")))
    inp = InputFile(os.path.join(GOLDEN, cfg["initial_input_file"]))
    inp.add_parameter("APER-MAP: x.txt")
    inp.add_parameter("SUMMARY-FILE: x"); inp.add_parameter("STAT-FILE: x"); inp.add_parameter("APER-FILE: x")
    inp.add_parameter("FLOW-FILE: x"); inp.add_parameter("PRESS-FILE: x"); inp.add_parameter("VTK-FILE: x")
    bulk = BulkRun(inp, **cfg["bulk_run_keyword_args"])
    bulk.generate_input_files(cfg["default_run_parameters"], cfg["default_file_formats"],
                              cfg["case_identifier"], cfg["case_parameters"], append=True)
    # 3 maps x 2 roughness x (3 + 3 + 2 outlet pressures) = 16 simulations
    assert len(bulk.input_file_list) == 16
    names = {f.outfile_name for f in bulk.input_file_list}
    assert "parallel-plate-01vox-RF0.00-500PA-INIT.INP" in names
    assert "Fracture1ApertureMap-10avg-RF1.00-200PA-INIT.INP" in names
    assert all(f["INLET-PRESS"].value == "1000" for f in bulk.input_file_list)


@pytest.mark.skipif(not os.path.isdir("/root/reference/apmapflow"), reason="reference tree not present (GPU box)")
def test_input_file_matches_reference_implementation():
    """Live diff against the reference's own InputFile in the authoring container."""
    sys.path.insert(0, "/root/reference")
    try:
        from apmapflow.run_model.run_model import InputFile as RefInputFile
    except Exception as exc:       # missing optional deps of the reference
        pytest.skip("reference package not importable: %s" % exc)
    finally:
        sys.path.remove("/root/reference")
    path = os.path.join(GOLDEN, "coverage_test.inp")
    ours, ref = InputFile(path), RefInputFile(path)
    assert list(ours.keys()) == list(ref.keys())
    for key in ours:
        assert (ours[key].value, ours[key].unit, ours[key].commented_out, ours[key].line) == \
               (ref[key].value, ref[key].unit, ref[key].commented_out, ref[key].line), key
    upd = {"OUTLET-SIDE": "LEFT", "apmap": "fracture-1", "INLET-PRESS": "250"}
    ours.update(upd)
    ref.update(upd)
    assert str(ours) == str(ref)
    assert ours.filename_format_args == ref.filename_format_args


def test_bulk_run_schedules_by_per_device_memory(monkeypatch):
    """``sys_RAM`` is the budget of EACH device, a case goes to the device with the most room, and
    never more than ``num_CPUs`` simulations are in flight (scheduling only: ``run_model`` is replaced by a stub)."""
    from apmapflow_b200.run_model import bulk_run as br

    class Handle:
        def __init__(self, inp, device):
            self.input_file, self.device, self.polls = inp, device, 0

        def poll(self):
            self.polls += 1
            return 0 if self.polls > 2 else None

    started, peak = [], {}
    in_flight = []

    def fake_run_model(inp, device=0, **kw):
        h = Handle(inp, device)
        started.append((inp.tag, device))
        in_flight[:] = [p for p in in_flight if p.polls <= 2] + [h]
        used = {}
        for p in in_flight:
            used[p.device] = used.get(p.device, 0.0) + p.input_file.RAM_req
        for d, u in used.items():
            peak[d] = max(peak.get(d, 0.0), u)
        peak["n"] = max(peak.get("n", 0), len(in_flight))
        return h

    monkeypatch.setattr(br, "run_model", fake_run_model)
    sizes = {"big": 6.0, "mid": 3.0, "small": 1.0}
    monkeypatch.setattr(br, "estimate_req_RAM", lambda maps, avail, **kw: [sizes[os.path.basename(m)] for m in maps])
    inp = InputFile(os.path.join(GOLDEN, "test-model-inputs.txt"))
    bulk = BulkRun(inp, num_CPUs=3, sys_RAM=8.0, devices=[0, 1], spawn_delay=0.0, retest_delay=0.0)
    for k, name in enumerate(["big", "big", "mid", "mid", "small", "small", "small", "big"]):
        case = inp.clone()
        case["APER-MAP"] = name
        case.tag = "%s-%d" % (name, k)
        bulk.input_file_list.append(case)
    bulk.start()
    assert sorted(t for t, _ in started) == sorted("%s-%d" % (n, k) for k, n in
                                                   enumerate(["big", "big", "mid", "mid", "small", "small", "small", "big"]))
    assert {d for _, d in started} == {0, 1}                      # both devices used
    assert peak[0] <= bulk.avail_RAM and peak[1] <= bulk.avail_RAM  # 7.2 GB each: two "big" never share a device
    assert peak["n"] <= 3
    assert started[0][1] != started[1][1]                          # equal first cases go to different devices
    with pytest.raises(EnvironmentError):                          # a map no device can hold is refused up front
        rm.estimate_req_RAM([os.path.join(GOLDEN, "parallel-plate-01vox.txt")], 1e-9)


@pytest.mark.skipif(not os.path.isdir("/root/reference/apmapflow"), reason="reference tree not present (GPU box)")
def test_bulk_run_input_files_match_reference_implementation():
    """Live diff against the reference's own BulkRun in the authoring container: the same YAML description must give
    the same list of input files -- order, file names and every line of their text (bulk_run.py:130-234)."""
    import yaml
    sys.path.insert(0, "/root/reference")
    try:
        from apmapflow.run_model import BulkRun as RefBulkRun, InputFile as RefInputFile
    except Exception as exc:
        pytest.skip("reference package not importable: %s" % exc)
    finally:
        sys.path.remove("/root/reference")
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "test-bulk-run.yaml")))
    path = os.path.join(GOLDEN, cfg["initial_input_file"])
    lists = []
    for bulk_cls, inp_cls in ((BulkRun, InputFile), (RefBulkRun, RefInputFile)):
        inp = inp_cls(path)
        for key in ("APER-MAP", "SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE"):
            inp[key].commented_out = False
        bulk = bulk_cls(inp, **cfg["bulk_run_keyword_args"])
        bulk.generate_input_files(cfg["default_run_parameters"], cfg["default_file_formats"],
                                  case_identifer=cfg["case_identifier"], case_params=cfg["case_parameters"])
        lists.append([(f.outfile_name, str(f)) for f in bulk.input_file_list])
        # a second description appended to the first (what apm_bulk_run does with several YAML files)
        bulk.generate_input_files({"OUTLET-PRESS": ["50"], "aper-map": ["parallel-plate-01vox"], "ROUGHNESS": ["0.00"]},
                                  cfg["default_file_formats"], append=True)
        lists.append([(f.outfile_name, str(f)) for f in bulk.input_file_list])
    assert lists[0] == lists[2] and len(lists[0]) == 16
    assert lists[1] == lists[3] and len(lists[1]) == 17


@pytest.mark.skipif(not os.path.isdir("/root/reference/apmapflow"), reason="reference tree not present (GPU box)")
def test_arg_input_differential_against_reference():
    """20 000 random input-file lines -- keywords with and without colons, values, units, quotes, comments, leading
    semicolons, odd spacing -- through both ArgInput implementations: same keyword, value, unit, comment, flag and the
    same line written back, also after changing the value and the unit."""
    import random
    sys.path.insert(0, "/root/reference")
    try:
        from apmapflow.run_model.run_model import ArgInput as RefArgInput
    except Exception as exc:
        pytest.skip("reference package not importable: %s" % exc)
    finally:
        sys.path.remove("/root/reference")
    rng = random.Random(20240601)
    keys = ["INLET-PRESS:", "OUTLET-SIDE:", "APER-MAP:", "MAP AVERAGING FACTOR:", "VOXEL SIZE:", "OVERWRITE", "EXE-FILE:",
            "ROUGHNESS REDUCTION:", "KEY", "K:", ":", "A:B:", "HIGH-MASK:", "OUTPUT-UNITS:", ""]
    vals = ["100", "1.0E-3", "TOP", '"a b.txt"', '"unterminated', "C:\\\\maps\\\\m.txt", "./rel/path.txt", "PA,MM,MM^3/MIN",
            "a=b", "10 MICRONS", "", "x y z", "'single'", "-5", "1e0", "\ttab"]
    units = ["PA", "KPA", "MICRONS", "CP", "", "KG/M^3", "unit with blanks"]
    comments = ["", ";comment", " ; spaced comment", ";;double", ";c1 ;c2", " ;"]

    def same(a, b, what, line):
        for attr in ("keyword", "value", "unit", "comment_msg", "commented_out", "line"):
            assert getattr(a, attr) == getattr(b, attr), (what, attr, line, getattr(a, attr), getattr(b, attr))

    compared = 0
    for _ in range(20000):
        parts = [rng.choice(["", ";", "; ", ";;"]) if rng.random() < 0.3 else "", rng.choice(keys)]
        if rng.random() < 0.85:
            parts.append(rng.choice(vals))
        if rng.random() < 0.5:
            parts.append(rng.choice(units))
        sep = rng.choice([" ", "  ", "\t", " "])
        line = parts[0] + sep.join(p for p in parts[1:]) + rng.choice(comments)
        if rng.random() < 0.1:
            line = "   " + line + "  "
        try:
            ref = RefArgInput(line)
        except Exception as exc:                       # whatever the reference rejects, we must reject too
            with pytest.raises(type(exc)):
                ArgInput(line)
            continue
        ours = ArgInput(line)
        same(ours, ref, "parsed", line)
        new_value = rng.choice(["42", "new value", ("7", True), ("8", False), '"q r"'])
        ours.value = new_value
        ref.value = new_value
        same(ours, ref, "value set", line)
        if ref.unit is not None:
            ours.unit = "PSI"
            ref.unit = "PSI"
            same(ours, ref, "unit set", line)
        compared += 1
    assert compared > 15000


@pytest.mark.skipif(not os.path.isdir("/root/reference/apmapflow"), reason="reference tree not present (GPU box)")
def test_input_file_differential_updates_clones_and_formats(tmp_path):
    """Random sequences of update / clone / add_parameter / filename formats on our InputFile and the reference's: the
    text written to disk, the output file name, the uncommented keys and the format arguments stay identical."""
    import random
    sys.path.insert(0, "/root/reference")
    try:
        from apmapflow.run_model.run_model import InputFile as RefInputFile
    except Exception as exc:
        pytest.skip("reference package not importable: %s" % exc)
    finally:
        sys.path.remove("/root/reference")
    rng = random.Random(7)

    def both(pair, fn):
        """Applies fn to ours and to the reference's; whatever the reference raises, ours must raise too."""
        out, errs = [], []
        for f in pair:
            try:
                out.append(fn(f))
                errs.append(None)
            except Exception as exc:
                out.append(None)
                errs.append(type(exc))
        assert errs[0] == errs[1], errs
        return out, errs[0]

    compared = 0
    for path in ("test-model-inputs.txt", "coverage_test.inp"):
        for trial in range(60):
            fmts = {"input_file": os.path.join(str(tmp_path), "{apmap}-{INLET-PRESS}-%d.inp" % trial),
                    "SUMMARY-FILE": "{apmap}-RF{ROUGHNESS}-log.txt", "VTK-FILE": "out/{apmap}.vtk"}
            if rng.random() < 0.2:
                fmts["NO-SUCH-FILE"] = "x.txt"                      # a format for a keyword the file does not have
            pair = [cls(os.path.join(GOLDEN, path), dict(fmts)) for cls in (InputFile, RefInputFile)]
            if rng.random() < 0.8:
                both(pair, lambda f: f.update({"apmap": "m0"}))
            for step in range(rng.randint(1, 6)):
                op = rng.choice(["update", "clone", "add", "comment", "str"])
                if op == "update":
                    upd = {rng.choice(["INLET-PRESS", "OUTLET-PRESS", "ROUGHNESS", "apmap", "OUTLET-SIDE", "MAP", "new-fmt-arg"]):
                           rng.choice(["1", "250.5", "TOP", "frac-7", ("9", True), ("3", False)]) for _ in range(rng.randint(1, 3))}
                    both(pair, lambda f: f.update(upd))
                elif op == "clone":
                    clones, err = both(pair, lambda f: f.clone())
                    if err is None:
                        pair = clones
                elif op == "add":
                    line = rng.choice([";EXTRA-KEY: 5 MM", "EXTRA-KEY: 6 CM ;note", "; a free comment line", "LOW-MASK: 0.5"])
                    both(pair, lambda f: f.add_parameter(line))
                elif op == "comment":
                    key = rng.choice(["OVERWRITE", "INLET-PRESS", "FLUID-VISCOSITY"])

                    def flip(f):
                        f[key].commented_out = not f[key].commented_out
                    both(pair, flip)
                else:
                    texts, _ = both(pair, str)
                    assert texts[0] == texts[1]
            texts, err = both(pair, str)
            assert texts[0] == texts[1]
            ours, ref = pair
            if err is None:
                assert ours.outfile_name == ref.outfile_name
                compared += 1
            assert list(ours.get_uncommented_values()) == list(ref.get_uncommented_values())
            assert ours.filename_format_args == ref.filename_format_args
    assert compared > 40
