"""apmapflow_b200 -- B200-native drop-in for the LCL hot path of stadelmanma/netl-ap-map-flow.

Mirrors the reference's ``apmapflow.run_model`` surface (InputFile, run_model, BulkRun,
estimate_req_RAM; apmapflow/run_model/__init__.py:18-20 and the names re-exported by
apmapflow/__init__.py) on top of libapmlcl.so.  The numerics are in the CUDA library; this package
is host glue.
"""
import logging
import os

__version__ = "0.1.0"

DEFAULT_MODEL_PATH = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_MODEL_NAME = "apm-lcl-model.exe"

_MAIN_LOGGER = "APM"


def _get_logger(name):
    """Logger in the ``APM.*`` tree like the reference (apmapflow/ap_map_flow.py:479-503)."""
    name = name.replace("apmapflow_b200", _MAIN_LOGGER).replace("apmapflow", _MAIN_LOGGER)
    if not name.startswith(_MAIN_LOGGER):
        name = _MAIN_LOGGER + "." + name
    return logging.getLogger(name)


def set_main_logger_level(level):
    if isinstance(level, str):
        level = getattr(logging, level.upper())
    root = logging.getLogger(_MAIN_LOGGER)
    if not root.handlers and not logging.getLogger().handlers:
        logging.basicConfig(format="%(levelname)s:%(name)s:%(message)s")
    root.setLevel(level)


from . import run_model  # noqa: E402,F401  (the subpackage, like `from apmapflow import run_model`)
