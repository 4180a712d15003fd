r"""
Reads one or more YAML bulk-run descriptions, builds the InputFiles of the sweep and runs them through the LCL
model, several at a time.  Without ``--start`` only a dry run is made (the input files are written for
inspection).  The keyword arguments of the FIRST YAML file instantiate the BulkRun.

Same command line and YAML keys as the reference's ``apm_bulk_run`` (apmapflow/scripts/apm_bulk_run.py:28-86:
``initial_input_file``, ``bulk_run_keyword_args``, ``default_run_parameters``, ``default_file_formats``,
``case_identifier``, ``case_parameters``) plus ``--devices`` (also accepted as a ``devices`` entry of
``bulk_run_keyword_args``).  The YAML is read with ``yaml.safe_load``: the reference's bare ``yaml.load(f)``
(:62) is a TypeError on PyYAML >= 6.
"""
import argparse
import os
import sys

import yaml

from .. import _get_logger, set_main_logger_level
from ..run_model import BulkRun, InputFile

logger = _get_logger("apmapflow.scripts")


def _device_list(text):
    return [int(d) for d in text.replace(",", " ").split()]


def build_parser():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("-v", "--verbose", action="store_true", help="debug messages are printed to the screen")
    parser.add_argument("--start", action="store_true", default=False,
                        help='flag must be supplied to "start" the bulk run')
    parser.add_argument("--devices", default=None, type=_device_list,
                        help="CUDA devices the cases are spread over, e.g. 0,1,2,3 (default: device 0)")
    parser.add_argument("--wait", action="store_true", default=False,
                        help="return only when the last simulations have finished (the interpreter waits for them at "
                             "exit either way)")
    parser.add_argument("input_files", nargs="+", type=os.path.realpath, help="1 or more YAML input files to load")
    return parser


def main(argv=None):
    set_main_logger_level("info")
    namespace = build_parser().parse_args(argv)
    if namespace.verbose:
        set_main_logger_level("debug")
    bulk_run = None
    logger.debug("Processing {} run parameter files".format(len(namespace.input_files)))
    for path in namespace.input_files:
        logger.debug("Reading {} parameter file.".format(os.path.basename(path)))
        with open(path, "r") as handle:
            inputs = yaml.safe_load(handle)
        inp_file = InputFile(inputs["initial_input_file"])
        if bulk_run is None:
            logger.debug("Instantiating initial BulkRun class")
            kwargs = dict(inputs.get("bulk_run_keyword_args") or {})
            if namespace.devices:
                kwargs["devices"] = namespace.devices
            bulk_run = BulkRun(inp_file, **kwargs)
        bulk_run.generate_input_files(inputs["default_run_parameters"],
                                      inputs.get("default_file_formats", None),
                                      case_identifer=inputs.get("case_identifier", None),
                                      case_params=inputs.get("case_parameters", None),
                                      append=True)
    if namespace.start:
        bulk_run.start()
        if namespace.wait:
            bulk_run.wait()
    else:
        bulk_run.dry_run()
        print('Add "--start" flag to begin simulations')
    return 0


if __name__ == "__main__":
    sys.exit(main())
