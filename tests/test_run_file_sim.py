"""File pipeline (csrc/host_io.cpp: .inp parser, unit tables, Fortran-format writers) checked against
the reference's golden output files.  The numeric core runs through the CPU emulation build of the
kernels (tests/sim) here; tests/test_gpu_run_file.py repeats it through libapmlcl.so on the GPU."""
import ctypes as C
import gzip
import os

import pytest

from sim import sim_lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_text(name):
    path = os.path.join(GOLDEN, name)
    if os.path.exists(path + ".gz"):
        return gzip.open(path + ".gz", "rt").read()
    return open(path).read()


def run(lib, workdir, inp_text, name="case.inp"):
    path = os.path.join(workdir, name)
    with open(path, "w") as f:
        f.write(inp_text)
    buf = C.create_string_buffer(1 << 22)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        rc = lib.apm_lcl_run_file(name.encode(), 0, buf, len(buf))
    finally:
        os.chdir(cwd)
    return rc, buf.value.decode()


def coverage_inp(extra=""):
    text = open(os.path.join(GOLDEN, "coverage_test.inp")).read()
    text = text.replace(';APER-MAP:   "(PATH/TO/FILE.*)"', "APER-MAP: " + os.path.join(GOLDEN, "ParallelPlate-1vox.txt"))
    return text + "\nBAD-INPUT: VALUE\n" + extra


def differing_lines(ours, gold):
    a, b = ours.split("\n"), gold.split("\n")
    assert len(a) == len(b)
    return [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x.rstrip() != y.rstrip()]


def check_against_golden_files(workdir):
    """What test/fortran/bin/test_model_code_coverage:115-155 diffs, at the precision an f64 solve can
    reach (the goldens come from the quad build)."""
    for ours, gold in [("coverage_test_pres.csv", "coverage_test_pres-orig.csv"),
                       ("coverage_test_aper.csv", "coverage_test_aper-orig.csv"),
                       ("coverage_test_flow-z.csv", "coverage_test_flow-z-orig.csv"),
                       ("coverage_test_flow-m.csv", "coverage_test_flow-m-orig.csv")]:
        assert differing_lines(open(os.path.join(workdir, ours)).read(), golden_text(gold)) == [], ours
    # stat files: only the APER-MAP path line and 12th-digit rounding of flow-derived values may differ
    for ours, gold in [("coverage_test_stats.csv", "coverage_test_stats-orig.csv"),
                       ("coverage_test_stats.yaml", "coverage_test_stats-orig.yaml")]:
        diff = differing_lines(open(os.path.join(workdir, ours)).read(), golden_text(gold))
        assert len(diff) <= 3 and diff[0][0] == 0, diff
        for _, x, y in diff[1:]:
            assert len(x) == len(y)
            assert sum(c != d for c, d in zip(x, y)) <= 2, (x, y)
    # VTK: header, points, APERTURE / PRESSURE / REYNOLDS sections identical; the residual and the
    # x component of the flow vectors are rounding noise in both builds
    sec, bad = "header", {}
    a = open(os.path.join(workdir, "coverage_test_data.vtk")).read().split("\n")
    b = golden_text("coverage_test_data-orig.vtk").split("\n")
    assert len(a) == len(b)
    for x, y in zip(a, b):
        if y.startswith(("SCALARS", "VECTORS", "POINTS")):
            sec = y.split()[1]
        if x.rstrip() != y.rstrip():
            bad[sec] = bad.get(sec, 0) + 1
            if sec.startswith("FLOW"):
                assert x.split()[1:] == y.split()[1:]       # y and z components identical
    assert set(bad) <= {"NORMALIZED_RESIDUAL[-]", "FLOW[M^3/SEC]"}, bad
    # log: same line count; solver name / time / residual-noise lines differ
    diff = differing_lines(open(os.path.join(workdir, "coverage_test_log.txt")).read(), golden_text("coverage_test_log-orig.txt"))
    allowed = ("USING", "REQUIRED TIME", "MAXIMUM NET FLOW RESIDUAL", "|")
    assert all(any(k in x for k in allowed) for _, x, _ in diff), diff


def test_coverage_case_reproduces_golden_files(tmp_path):
    rc, out = run(sim_lib.load(), str(tmp_path), coverage_inp(), "coverage_test.inp")
    assert rc == 0, out
    assert "USING SUPPLIED INPUT FILE coverage_test.inp" in out and "**** EXITING PROGRAM ****" in out
    assert "** WARNING: UNRECOGNIZED INPUT: BAD-INPUT: VALUE **" in out
    check_against_golden_files(str(tmp_path))


def plate(tmp_path, value, nz=20, nx=16):
    """A small parallel-plate map file (the 100 x 100 fixtures cost ~15 s per run in the CPU emulation)."""
    path = os.path.join(str(tmp_path), "plate-%gvox.txt" % value)
    with open(path, "w") as f:
        f.write("\n".join(" ".join(["%.1f" % value] * nx) for _ in range(nz)) + "\n")
    return path


def small_inp(tmp_path, body, amap="BadApertureMap.txt"):
    return ("APER-MAP: %s\nSUMMARY-FILE: %s/log.txt\nSTAT-FILE: %s/stat.csv\nAPER-FILE: %s/aper.csv\n"
            "FLOW-FILE: %s/flow.csv\nPRESS-FILE: %s/pres.csv\nVTK-FILE: %s/data.VTK\nOVERWRITE\n"
            % ((os.path.join(GOLDEN, amap),) + (str(tmp_path),) * 6)) + body


OK_BODY = ("INLET-PRESS: 100 PA\nOUTLET-PRESS: 0 PA\nOUTLET-SIDE: TOP\nFLUID-DENSITY: 1000 KG/M^3\n"
           "FLUID-VISCOSITY: 1.0 CP\nMAP AVERAGING FACTOR: 1\nVOXEL SIZE: 10 MICRONS\nLOW-MASK: 0.0001\n")


@pytest.mark.parametrize("body,amap,pattern", [      # test_model_code_coverage:214-272
    (OK_BODY, "BadApertureMap.txt", r"FIELD\s*\d+\s*ON THE"),
    (OK_BODY.replace("FLUID-VISCOSITY: 1.0 CP", "FLUID-VISCOSITY: -1.0 CP"), "parallel-plate-01vox.txt",
     "ERROR - NONE OR INVAILD VISCOSITY"),
    (OK_BODY.replace("OUTLET-SIDE: TOP", "OUTLET-SIDE: NOWHERE"), "PLATE",      # detected after MAP_STATS: small map
     "INVALID OUTLET SIDE ENTERED: NOWHERE"),
    (OK_BODY + "MAXIMUM MAP DIMENSION: 50\n", "parallel-plate-01vox.txt", "MAXIMUM NUMBER OF HORIZONTAL"),
    (OK_BODY.replace("INLET-PRESS: 100 PA", "INLET-PRESS: 100"), "parallel-plate-01vox.txt", "INSUFFICIENT"),
    (OK_BODY.replace("INLET-PRESS: 100 PA", "INLET-PRESS: abc PA"), "parallel-plate-01vox.txt", r"FIELD\s*2\s*ON THE"),
])
def test_failure_cases_return_1_with_reference_messages(tmp_path, body, amap, pattern):
    import re
    if amap == "PLATE":
        amap = plate(tmp_path, 1.0)
    rc, out = run(sim_lib.load(), str(tmp_path), small_inp(tmp_path, body, amap))
    assert rc == 1
    assert re.search(pattern, out), out


def test_missing_input_and_map_and_overwrite(tmp_path):
    lib = sim_lib.load()
    buf = C.create_string_buffer(4096)
    assert lib.apm_lcl_run_file(b"", 0, buf, 4096) == 1
    assert b"FATAL ERROR - NO INPUT FILE" in buf.value
    rc, out = run(lib, str(tmp_path), "OVERWRITE\n" + OK_BODY)
    assert rc == 1 and "FATAL ERROR - NO APERTURE MAP" in out
    # second run without OVERWRITE refuses to clobber and removes nothing that existed
    amap = plate(tmp_path, 1.0)
    rc, out = run(lib, str(tmp_path), small_inp(tmp_path, OK_BODY, amap))
    assert rc == 0, out
    assert os.path.exists(tmp_path / "data.vtk")          # extension lower-cased (IO_MODULE.F:169-170)
    assert os.path.exists(tmp_path / "flow-x.csv") and os.path.exists(tmp_path / "stat.yaml")
    rc, out = run(lib, str(tmp_path), small_inp(tmp_path, OK_BODY, amap).replace("OVERWRITE\n", ""))
    assert rc == 1 and 'SPECIFY "OVERWRITE" IN THE' in out
    assert os.path.exists(tmp_path / "pres.csv")


def test_partially_failing_open_io_removes_what_it_created(tmp_path):
    """IO_MODULE.F:219-226: when any output file cannot be opened, every file OPEN_IO did open is removed
    again -- also when the log file itself opened fine.  Left behind, the empty files would make the next
    run (after the user removed the one offender) fail on each of them in turn."""
    lib = sim_lib.load()
    (tmp_path / "pres.csv").write_text("keep me\n")
    inp = small_inp(tmp_path, OK_BODY, plate(tmp_path, 1.0)).replace("OVERWRITE\n", "")
    rc, out = run(lib, str(tmp_path), inp)
    assert rc == 1 and "pres.csv ALREADY EXISTS" in out
    left = sorted(f for f in os.listdir(tmp_path) if not f.endswith(".inp") and not f.startswith("plate-"))
    assert left == ["pres.csv"], left
    assert (tmp_path / "pres.csv").read_text() == "keep me\n"
    os.remove(tmp_path / "pres.csv")
    rc, out = run(lib, str(tmp_path), inp)
    assert rc == 0, out


def test_units_and_rate_bc_through_files(tmp_path):
    body = ("INLET-RATE: 1.0 MM^3/SEC\nOUTLET-PRESS: 1 KPA\nOUTLET-SIDE: RIGHT\nFLUID-DENSITY: 1 G/CM^3\n"
            "FLUID-VISCOSITY: 1.0 CP\nMAP AVERAGING FACTOR: 10\nVOXEL SIZE: 10 MICRONS\nLOW-MASK: 0.0001\n"
            "OUTPUT-UNITS: KPA,MM,MM^3/SEC\n")
    nz, nx = 20, 16
    rc, out = run(sim_lib.load(), str(tmp_path), small_inp(tmp_path, body, plate(tmp_path, 10.0, nz, nx)))
    assert rc == 0, out
    assert "FRACTURE FLOW IS RATE CONTROLLED" in out
    import yaml
    st = yaml.safe_load(open(tmp_path / "stat.yaml"))
    assert st["OUTLET RATE"][1] == "MM^3/SEC" and st["OUTLET RATE"][0] == pytest.approx(1.0, rel=1e-6)
    assert st["OUTLET PRESS"] == [1.0, "KPA"]
    assert st["INLET PRESS"][0] > 1.0
    assert st["DENSITY"] == [1.0, "G/CM^3"]
    # analytic cubic law for a 10-voxel parallel plate gives the inlet pressure
    w, l = nz * 10e-6 * 10, nx * 10e-6 * 10        # outlet RIGHT: the flow runs along x
    b = 10 * 10e-6
    dp = 1e-9 * 12 * 1e-3 * l / (w * b ** 3)
    assert (st["INLET PRESS"][0] - 1.0) * 1e3 == pytest.approx(dp, rel=1e-6)


def test_map_reader_notations_and_separators(tmp_path):
    """READ_DATA_MAP through LNPROC accepts blanks, commas, TABs and '=' as separators and every Fortran spelling of a
    real (MAP_MODULE.F:18-85, STRING_MODULE.F:114-239).  The reader's fast path (std::from_chars on plain decimal
    fields) and its Fortran-minded path (D exponents, leading '+', ...) must agree: the same values written in different
    notations give byte-identical output files; a field that is no number fails with the reference's message."""
    lib = sim_lib.load()
    rows = [[1.0, 2.5, 0.5, 3.0, 12.0, 7.25], [4.0, 1.5, 2.0, 0.25, 9.0, 1.0], [2.0, 2.0, 6.5, 1.0, 3.0, 8.0],
            [5.0, 0.75, 1.0, 2.0, 4.0, 4.0]]
    plain = "\n".join(" ".join("%.6f" % v for v in row) for row in rows) + "\n"
    spell = {1.0: ["1", "1.", "1e0", "1E0", "+1.0", "1d0", "1.0D+00", "0.1e1", "10.E-1"], 2.5: ["2.5", "+2.5", ".25e1", "2.5D0", "25e-1"],
             0.5: [".5", "0.5", "5.E-1", "+.5", "5d-1"], 0.25: [".25", "2.5e-1"], 0.75: ["0.75", "7.5D-1"]}
    out = []
    for z, row in enumerate(rows):
        toks = []
        for x, v in enumerate(row):
            forms = spell.get(v, ["%g" % v, "%.3f" % v, "%.3E" % v])
            toks.append(forms[(z * 7 + x) % len(forms)])
        sep = [" ", ",", "\t", "  ,  ", " = "][z % 5]
        out.append(sep.join(toks) + ("   ; a comment" if z == 1 else ""))
    mixed = "\r\n".join(out) + "\r\n\n"
    results = []
    for name, text in (("plain", plain), ("mixed", mixed)):
        work = tmp_path / name
        work.mkdir()
        (work / "map.txt").write_text(text)
        rc, log = run(lib, str(work), small_inp(work, OK_BODY, str(work / "map.txt")))
        assert rc == 0, log
        results.append({f: (work / f).read_text() for f in ("aper.csv", "pres.csv", "flow-z.csv", "flow-x.csv")})
    assert results[0] == results[1]
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "map.txt").write_text("1.0 2.0 3.0\n1.0 2..0 3.0\n")
    rc, log = run(lib, str(bad), small_inp(bad, OK_BODY, str(bad / "map.txt")))
    import re
    assert rc == 1 and re.search(r"FIELD\s*2\s*ON THE", log), log
    for token in ("1e", "0x10", "nan", "--1", "1e+"):
        (bad / "map.txt").write_text("1.0 2.0 3.0\n1.0 %s 3.0\n" % token)
        rc, log = run(lib, str(bad), small_inp(bad, OK_BODY, str(bad / "map.txt")))
        assert rc == 1 and re.search(r"FIELD\s*2\s*ON THE", log), (token, log)


def test_parallel_writers_and_reader_equal_the_sequential_ones(tmp_path, monkeypatch):
    """APM_LCL_IO_THREADS=1 takes the sequential reader and writers; any other count must give the same bytes in every
    output file (rows are formatted by a pool of threads and written in order)."""
    lib = sim_lib.load()
    outs = {}
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("APM_LCL_IO_THREADS", threads)
        work = tmp_path / ("t" + threads)
        work.mkdir()
        rc, log = run(lib, str(work), coverage_inp(), "coverage_test.inp")
        assert rc == 0, log
        outs[threads] = {f: (work / f).read_bytes() for f in sorted(os.listdir(work)) if f != "coverage_test_log.txt"}
        assert len(outs[threads]) >= 9
    assert outs["1"] == outs["3"] == outs["16"]
