"""Host side of the drop-in boundary: the reference's ``apmapflow.run_model.run_model`` API
(apmapflow/run_model/run_model.py) re-pointed at libapmlcl.so.

What is kept: ``ArgInput`` (:28), ``InputFile`` (:274) with parse / clone / update / filename formats /
``;EXE-FILE:`` handling, ``estimate_req_RAM`` (:657) and ``run_model`` (:710) with its argument
meaning and its Popen-like return value (``poll()``, ``wait()``, ``returncode``, ``pid``,
``input_file``, ``start_time``, and after completion ``stdout_content``, ``stderr_content``,
``end_time``).  What changes: by default the model does not run in a child process, the call goes
through ctypes into ``apm_lcl_run_file`` on a worker thread (ctypes drops the GIL); a custom
``;EXE-FILE:`` is still honoured with a real ``Popen``.
"""
import ctypes as C
import os
import re
import shlex
import threading
import time
from collections import OrderedDict
from subprocess import PIPE, Popen

from .. import _get_logger

logger = _get_logger(__name__)
DEFAULT_MODEL_PATH = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DEFAULT_MODEL_NAME = "apm-lcl-model.exe"
_LOG_CAP = 1 << 22


class ArgInput(object):
    """One line of an LCL input file: ``KEYWORD: value unit ;comment``.

    A leading ``;`` marks the whole parameter as commented out, text after a later ``;`` is kept as
    ``comment_msg``; without a ``KEYWORD:`` field the whole line is the value
    (reference behaviour pinned by test/unit/TestRunCore.py:23-72)."""

    def __init__(self, line):
        self._fields = []
        self._ivalue = -1
        self.comment_msg = ""
        self.commented_out = False
        self._parse(line)

    def _parse(self, line):
        text = line.strip()
        if text.startswith(";"):
            text = text[1:]
            self.commented_out = True
        cut = text.find(";")
        if cut >= 0:
            self.comment_msg = text[cut:].strip()
            text = text[:cut].strip()
        try:
            fields = shlex.split(text, posix=False)      # keeps the double quotes and windows paths
        except ValueError:
            fields = text.split()
        self._fields = fields or [""]
        self._ivalue = -1
        if re.search(r":(:?\s|$)", text):
            for k, fld in enumerate(self._fields):
                if fld.endswith(":"):
                    self._ivalue = k + 1
                    while len(self._fields) <= self._ivalue + 1:
                        self._fields.append("")           # slots for value and unit
                    break

    def __str__(self):
        return self.line

    @property
    def keyword(self):
        return re.sub(r":$", "", self._fields[0])

    @property
    def value(self):
        if self._ivalue > -1:
            return self._fields[self._ivalue]
        return " ".join(self._fields)

    @value.setter
    def value(self, value):
        commented = False
        if isinstance(value, (list, tuple)):
            value, commented = value[0], bool(value[1])
        self.commented_out = commented
        if self._ivalue == -1:
            keep = self.commented_out
            self._parse(str(value))
            self.commented_out = keep
        else:
            self._fields[self._ivalue] = str(value)

    @property
    def unit(self):
        return self._fields[self._ivalue + 1] if self._ivalue > -1 else None

    @unit.setter
    def unit(self, unit):
        self._fields[self._ivalue + 1] = str(unit)

    @property
    def line(self):
        return (";" if self.commented_out else "") + " ".join(self._fields) + self.comment_msg


class InputFile(OrderedDict):
    """Reads, edits and writes LCL model input files; every value is an ``ArgInput``
    (apmapflow/run_model/run_model.py:274-655)."""

    def __init__(self, infile, filename_formats=None):
        super().__init__()
        self.filename_format_args = {}
        self.RAM_req = None
        self.filename_formats = dict(filename_formats or {})
        self.outfile_name = infile.outfile_name if isinstance(infile, InputFile) else infile
        self.executable = None
        self.parse_input_file(infile)
        if "input_file" not in (filename_formats or {}):
            self.filename_formats["input_file"] = self.outfile_name

    def __str__(self):
        self._construct_file_names()
        return "".join(arg.line + "\n" for arg in self.values())

    def __setitem__(self, key, value, new_param=False):
        if new_param:
            super().__setitem__(key, value)
        else:
            try:
                self[key].value = value
            except KeyError:
                raise KeyError("'{}' is not set, use .add_parameter method to set param".format(key))
        if key == "EXE-FILE":
            self.set_executable()

    def parse_input_file(self, infile):
        if isinstance(infile, InputFile):
            content = str(infile)
            self.infile = os.path.realpath(infile.outfile_name)
        else:
            self.infile = os.path.realpath(infile)
            with open(infile, "r") as handle:
                content = handle.read()
        for line in content.split("\n"):
            self.add_parameter(line)
        self.set_executable()

    def add_parameter(self, line):
        arg = ArgInput(re.sub(r"^(;+)\s+", r"\1", line))
        self.__setitem__(arg.keyword, arg, new_param=True)

    def set_executable(self):
        """``;EXE-FILE:`` (relative to the input file) selects another model executable; otherwise the
        packaged front end is used (reference :453-491)."""
        self.executable = None
        if "EXE-FILE" in self:
            self["EXE-FILE"].commented_out = True
            exe = self["EXE-FILE"].value
            if not os.path.isabs(exe):
                exe = os.path.realpath(os.path.join(os.path.dirname(self.infile), exe))
            if os.path.exists(exe):
                self.executable = exe
            else:
                logger.warning("The exe file specified does not exist: " + exe)
        if not self.executable:
            self.executable = os.path.join(DEFAULT_MODEL_PATH, DEFAULT_MODEL_NAME)

    def clone(self, file_formats=None):
        return InputFile(self, filename_formats=self.filename_formats if file_formats is None else file_formats)

    def update(self, *args, **kwargs):
        if len(args) > 1:
            raise TypeError("update expected at most 1 arguments, got {:d}".format(len(args)))
        for key, value in dict(*args, **kwargs).items():
            try:
                self[key] = value
            except KeyError:
                self.filename_format_args[key] = value
        self._construct_file_names()

    def get_uncommented_values(self):
        return OrderedDict((k, a) for k, a in self.items() if not a.commented_out)

    def _construct_file_names(self, make_dirs=False):
        fmt_args = {key: arg.value for key, arg in self.items()}
        fmt_args.update(self.filename_format_args)
        names = {key: fmt.format(**fmt_args) for key, fmt in self.filename_formats.items()}
        for key, name in names.items():
            try:
                self[key] = name
            except KeyError:
                if key != "input_file":
                    logger.error("Outfile: {} not defined in initialization file".format(key))
                    raise KeyError(key)
            if make_dirs:
                folder = os.path.split(name)[0] or "."
                if not os.path.isdir(folder):
                    os.makedirs(folder)
        self.outfile_name = names["input_file"]

    def write_inp_file(self, alt_path=None):
        self._construct_file_names(make_dirs=True)
        content = str(self)
        name = os.path.join(alt_path, self.outfile_name) if alt_path else self.outfile_name
        with open(name, "w") as handle:
            handle.write(content)
        logger.info("Input file saved as: " + name)


def map_dimensions(path, delim="auto"):
    """(nx, nz) of a text data map without converting its values: records are counted (blank and comment-only lines
    do not count) and only the LAST record is split into fields -- its width is the one the model uses
    (MAP_MODULE.F:18-85).  A map is read at disk speed; tokenising every line in Python took longer than the solve."""
    nz, last = 0, None
    with open(path, "rb") as handle:
        for line in handle:
            data = line.split(b";", 1)[0].strip()
            if not data:
                continue
            nz += 1
            last = data
    nx = len([t for t in re.split(rb"[\s,=]+", last) if t]) if last else 0
    return nx, nz


# device memory per cell: 9 fine f64 fields + interpolation weights + coarse levels (~1/3) + slack (measured:
# 43 GB resident for a 16384^2 map).  The CUDA context (~0.6 GB) is not part of it: runs are threads of this
# process and share one context per device, and a per-run constant would reject the reference's own bulk-run
# descriptions (test/fixtures/test-bulk-run.yaml budgets 0.01 GB for 100 x 100 maps).
_BYTES_PER_CELL = 160.0


def estimate_req_RAM(input_maps, avail_RAM=float("inf"), suppress=False, **kwargs):
    """Memory estimate per map in GB.  Same signature and error behaviour as the reference (:657-707),
    but the model is the O(N) device footprint of the multigrid solver, not the reference's
    0.00505193*(nx*nz)^1.4516 kB regression of the direct solver's fill."""
    ram_per_map = []
    error = False
    for fname in input_maps:
        nx, nz = map_dimensions(fname, kwargs.get("delim", "auto"))
        ram = _BYTES_PER_CELL * nx * nz / 2.0 ** 30
        ram_per_map.append(ram)
        if ram > avail_RAM:
            error = True
            logger.fatal("Map {} requires {} GBs of RAM only {} GBs was alloted.".format(fname, ram, avail_RAM))
    if error and not suppress:
        raise EnvironmentError
    return ram_per_map


class LclProcess(object):
    """Popen-compatible handle of an in-process model run (what ``run_model`` returns)."""

    def __init__(self, input_file, device, show_stdout, devices=None):
        self.devices = [int(d) for d in devices] if devices else None
        self.input_file = input_file
        self.args = ("libapmlcl.so:apm_lcl_run_file", input_file.outfile_name)
        self.pid = os.getpid()
        self.device = device
        self.returncode = None
        self.stdout_content = None
        self.stderr_content = None
        self.start_time = time.time()
        self.end_time = None
        self._show = show_stdout
        # NOT a daemon thread: the reference's runs are child processes that outlive the caller
        # (apm_bulk_run.py:81 returns right after start()); the interpreter joins this thread at exit, so a
        # script that ends while simulations are in flight still gets their output files complete
        self._thread = threading.Thread(target=self._run, daemon=False)
        self._thread.start()

    def _run(self):
        from .. import _lib
        try:
            lib = _lib.load()
            buf = C.create_string_buffer(_LOG_CAP)
            name = os.fsencode(self.input_file.outfile_name)
            if self.devices and len(self.devices) > 1:      # one map cut into row strips over these GPUs
                devs = (C.c_int * len(self.devices))(*self.devices)
                rc = lib.apm_lcl_run_file_striped(name, devs, len(self.devices), buf, _LOG_CAP)
            else:
                rc = lib.apm_lcl_run_file(name, int(self.device), buf, _LOG_CAP)
            out, err = buf.value.decode("utf-8", "replace"), ""
        except Exception as exc:   # missing library etc.: fail loudly through the handle
            rc, out, err = 1, "", "{}: {}".format(type(exc).__name__, exc)
        self.end_time = time.time()
        if self._show:
            print(out, end="")
            if err:
                print(err)
            self.stdout_content, self.stderr_content = None, None
        else:
            self.stdout_content, self.stderr_content = out, err
        minutes = (self.end_time - self.start_time) / 60.0
        logger.info("Completed Simulation:\n\tinput file: {}\n\tTime Required: {:0.3f} minutes\n\tExit Code: {}"
                    .format(self.input_file.outfile_name, minutes, rc))
        self.returncode = rc

    def poll(self):
        return self.returncode

    def wait(self, timeout=None):
        self._thread.join(timeout)
        if self.returncode is None:
            raise TimeoutError("model still running")
        return self.returncode

    def communicate(self, input=None, timeout=None):
        self.wait(timeout)
        return self.stdout_content, self.stderr_content

    def kill(self):          # an in-flight CUDA solve cannot be interrupted; it is short
        pass

    terminate = kill


class _AsyncCommunicate(threading.Thread):
    """Collects the output of a real child process (custom ``;EXE-FILE:``), reference :237-271."""

    def __init__(self, proc):
        super().__init__(daemon=True)
        self.proc = proc

    def run(self):
        out, err = self.proc.communicate()
        self.proc.stdout_content, self.proc.stderr_content = out, err
        self.proc.end_time = time.time()


def _lcl_library_setup():
    from .. import _lib
    lib = _lib.load()
    lib.apm_lcl_run_file.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    lib.apm_lcl_run_file.restype = C.c_int
    return lib


def run_model(input_file_obj, synchronous=False, show_stdout=False, device=0, devices=None):
    """Runs one LCL simulation described by an ``InputFile`` (reference :710-772).

    The input file is (re)written to ``input_file_obj.outfile_name`` first, exactly like the
    reference.  Returns a Popen-like handle; with ``synchronous=True`` it has finished on return.
    ``device`` (new) selects the GPU; ``devices`` (new) names several GPUs of this node to cut one large map
    into row strips over (one solver per GPU inside this process)."""
    input_file_obj.write_inp_file()
    default_exe = os.path.join(DEFAULT_MODEL_PATH, DEFAULT_MODEL_NAME)
    exe = os.path.abspath(input_file_obj.executable or default_exe)
    use_exe = os.path.exists(exe) and os.path.realpath(exe) != os.path.realpath(default_exe)
    if use_exe or os.environ.get("APM_LCL_FORCE_EXE"):
        logger.debug("Using executable located at: " + exe)
        out = None if show_stdout else PIPE
        env = dict(os.environ, APM_LCL_DEVICE=str(device))
        if devices:
            env["APM_LCL_DEVICES"] = ",".join(str(int(d)) for d in devices)
        proc = Popen((exe, input_file_obj.outfile_name), stdout=out, stderr=out, universal_newlines=True, env=env)
        proc.input_file = input_file_obj
        proc.start_time = time.time()
        proc.device = device
        comm = _AsyncCommunicate(proc)
        comm.start()
        if synchronous:
            comm.join()
        return proc
    _lcl_library_setup()          # raises OSError when libapmlcl.so is missing: no CPU fallback
    proc = LclProcess(input_file_obj, devices[0] if devices else device, show_stdout, devices=devices)
    logger.info("Beginning Simulation:\n\tInput File: {} \n\tProcess ID: {}".format(input_file_obj.outfile_name, proc.pid))
    if synchronous:
        proc.wait()
    return proc
