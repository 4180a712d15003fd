r"""
Runs each of the given LCL initialization files through the model, one after the other, with the model's
output on the screen.  A line ``;EXE-FILE: (exec file)`` in an input file, or ``-e``, selects another model
executable; without either the solve goes straight into libapmlcl.so on the chosen GPU.

Same command line as the reference's ``apm_run_lcl_model`` (apmapflow/scripts/apm_run_lcl_model.py:26-66:
``-v``, ``-e EXECUTABLE``, ``input_files...``) plus ``--device`` / ``--devices`` for the GPU(s).
"""
import argparse
import os
import sys

from .. import _get_logger, set_main_logger_level
from ..run_model import InputFile, estimate_req_RAM, run_model

logger = _get_logger("apmapflow.scripts")


def _device_list(text):
    return [int(d) for d in text.replace(",", " ").split()]


def build_parser():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("-v", "--verbose", action="store_true", help="debug messages are printed to the screen")
    parser.add_argument("-e", "--executable", default=None, type=os.path.realpath,
                        help="model executable to use instead of the in-process library call")
    parser.add_argument("-d", "--device", default=0, type=int, help="CUDA device of the solve (default 0)")
    parser.add_argument("--devices", default=None, type=_device_list,
                        help="several devices, e.g. 0,1,2,3: one large map is cut into row strips over them")
    parser.add_argument("input_files", nargs="+", type=os.path.realpath,
                        help="1 or more model parameter files to load")
    return parser


def main(argv=None):
    """Parses the command line and runs the files in order; returns the number of failed runs (the reference
    returns nothing and ignores the exit codes; a console entry point turns this into the exit status)."""
    set_main_logger_level("info")
    args = build_parser().parse_args(argv)
    if args.verbose:
        set_main_logger_level("debug")
    failed = 0
    for n, path in enumerate(args.input_files):
        logger.debug("Processing input file {} of {}: {}".format(n + 1, len(args.input_files), os.path.basename(path)))
        input_file = InputFile(path)
        ram_req = estimate_req_RAM([input_file["APER-MAP"].value], suppress=True)[0]
        logger.info("Map will require approximately {:0.6f} GBs of device memory".format(ram_req))
        if args.executable:
            input_file.executable = args.executable
        proc = run_model(input_file, synchronous=True, show_stdout=True, device=args.device, devices=args.devices)
        if proc.returncode:
            failed += 1
    return failed


if __name__ == "__main__":
    sys.exit(main())
