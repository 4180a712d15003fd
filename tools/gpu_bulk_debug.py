import os, sys, tempfile, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from apmapflow_b200.run_model import BulkRun, InputFile
GOLDEN = os.path.join(ROOT, "tests", "golden")
tmp = tempfile.mkdtemp()
cfg = yaml.safe_load(open(os.path.join(GOLDEN, "test-bulk-run.yaml")))
inp = InputFile(os.path.join(GOLDEN, cfg["initial_input_file"]))
for key in ("APER-MAP", "SUMMARY-FILE", "STAT-FILE", "APER-FILE", "FLOW-FILE", "PRESS-FILE", "VTK-FILE"):
    inp[key].commented_out = False
fmts = {k: os.path.join(tmp, v) for k, v in cfg["default_file_formats"].items()}
fmts["APER-MAP"] = os.path.join(GOLDEN, "{aper-map}.txt")
ncpu = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bulk = BulkRun(inp, **dict(cfg["bulk_run_keyword_args"], spawn_delay=0.0, retest_delay=0.01, sys_RAM=8.0, num_CPUs=ncpu))
bulk.generate_input_files(cfg["default_run_parameters"], fmts, cfg["case_identifier"], cfg["case_parameters"])
import importlib
br = importlib.import_module("apmapflow_b200.run_model.bulk_run")
procs = []
orig = br.run_model
def wrapped(*a, **k):
    p = orig(*a, **k); procs.append(p); return p
br.run_model = wrapped
bulk.start(); bulk.wait()
for p in procs:
    print(p.input_file.outfile_name.split('/')[-1], p.returncode)
    if p.returncode != 0:
        print(p.stdout_content[-1500:], p.stderr_content)
