"""the reference's own loaders must parse the files our writers produce
(auto-delimiter regex of DataField, apmapflow/ap_map_flow.py:69-87; YAML stat files read by
apm_combine_yaml_stat_files.py:92-96).  Needs /root/reference, so it only runs in the authoring
container; the numeric core goes through the CPU emulation build."""
import os
import sys

import numpy as np
import pytest

from sim import sim_lib
from test_run_file_sim import GOLDEN, coverage_inp, run

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/apmapflow"), reason="reference tree not present")


def _reference_module(name):
    sys.path.insert(0, "/root/reference")
    try:
        return __import__(name, fromlist=["x"])
    except Exception as exc:
        pytest.skip("reference module %s not importable: %s" % (name, exc))
    finally:
        sys.path.remove("/root/reference")


@pytest.fixture(scope="module")
def coverage_run(tmp_path_factory):
    work = tmp_path_factory.mktemp("downstream")
    rc, out = run(sim_lib.load(), str(work), coverage_inp(), "coverage_test.inp")
    assert rc == 0, out
    return work


def test_reference_datafield_parses_our_csv_maps(coverage_run):
    tmp_path = coverage_run
    DataField = _reference_module("apmapflow.ap_map_flow").DataField
    for name, expect in (("coverage_test_pres.csv", None), ("coverage_test_aper.csv", 1.0e-5),
                         ("coverage_test_flow-z.csv", None), ("coverage_test_flow-m.csv", None)):
        fld = DataField(str(tmp_path / name))
        assert (fld.nx, fld.nz) == (100, 100)
        assert np.isfinite(fld.data_map).all()
        if expect is not None:
            assert np.allclose(fld.data_map, expect)
    pres = DataField(str(tmp_path / "coverage_test_pres.csv")).data_map
    assert pres[0, 0] == pytest.approx(99.5) and pres[-1, 0] == pytest.approx(0.5)
    import yaml
    st = yaml.safe_load(open(tmp_path / "coverage_test_stats.yaml"))
    assert st["NX"] == 100 and st["OUTLET RATE"][1] == "M^3/SEC"


def test_reference_statfile_parser_reads_our_csv_stat_file(coverage_run):
    """The legacy CSV stat file (key line / value line blocks, units in brackets) through the reference's own parser
    (apm_convert_csv_stats_file.py:46-100) must give what our YAML stat file says."""
    import yaml
    mod = _reference_module("apmapflow.scripts.apm_convert_csv_stats_file")
    parsed = mod.StatFile(str(coverage_run / "coverage_test_stats.csv"))
    ours = yaml.safe_load(open(coverage_run / "coverage_test_stats.yaml"))
    assert parsed["NX"] == ours["NX"] == 100 and parsed["NZ"] == ours["NZ"] == 100
    assert os.path.basename(parsed.map_file) == "ParallelPlate-1vox.txt"
    checked = 0
    for key, value in ours.items():
        if isinstance(value, list) and key in parsed:
            assert parsed[key][0] == pytest.approx(value[0], rel=1e-11), key
            assert parsed[key][1] == value[1], key
            checked += 1
    assert checked >= 20
    assert parsed["OUTLET RATE"][0] == pytest.approx(0.833333333333e-14, rel=1e-9)


def test_reference_yaml_combiner_accepts_our_stat_files(coverage_run, tmp_path, monkeypatch):
    """apm_combine_yaml_stat_files.py:92-160: key order from the text of the first file, units hoisted into the
    header when they agree.  (Its bare yaml.load(f) is a TypeError on PyYAML >= 6, so the loader is pinned here.)"""
    import shutil
    import yaml
    mod = _reference_module("apmapflow.scripts.apm_combine_yaml_stat_files")
    real_load = yaml.load
    monkeypatch.setattr(mod.yaml, "load", lambda stream, Loader=yaml.SafeLoader: real_load(stream, Loader))
    first = str(coverage_run / "coverage_test_stats.yaml")
    second = str(tmp_path / "second_stats.yaml")
    shutil.copy(first, second)
    order = mod.determine_key_order(first)
    text_keys = [line.split(":")[0] for line in open(first).read().split("\n") if line and not line.startswith(" ")]
    assert order == [k for k in text_keys if k in order] and len(order) >= 25
    data = []
    for path in (first, second):
        data.append(yaml.safe_load(open(path)))
        data[-1]["_stat_file"] = path
    out = str(tmp_path / "combined.csv")
    mod.output_stat_data(out, order, data)
    lines = open(out).read().split("\n")
    assert len(lines) == 3 and lines[0].startswith("STAT-FILE-PATH,STAT-FILE,")
    assert "OUTLET RATE [M^3/SEC]" in lines[0]
    assert lines[1].split(",")[2:] == lines[2].split(",")[2:]
