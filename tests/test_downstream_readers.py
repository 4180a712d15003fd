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


def test_reference_datafield_parses_our_csv_maps(tmp_path):
    rc, out = run(sim_lib.load(), str(tmp_path), coverage_inp(), "coverage_test.inp")
    assert rc == 0, out
    sys.path.insert(0, "/root/reference")
    try:
        from apmapflow.ap_map_flow import DataField
    except Exception as exc:
        pytest.skip("reference package not importable: %s" % exc)
    finally:
        sys.path.remove("/root/reference")
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
