"""Run-model API, same names as apmapflow/run_model/__init__.py:18-20."""
from .run_model import DEFAULT_MODEL_NAME, DEFAULT_MODEL_PATH, ArgInput, InputFile, LclProcess
from .run_model import estimate_req_RAM, map_dimensions, run_model
from .bulk_run import BulkRun

__all__ = ["ArgInput", "InputFile", "LclProcess", "estimate_req_RAM", "map_dimensions", "run_model", "BulkRun",
           "DEFAULT_MODEL_NAME", "DEFAULT_MODEL_PATH"]
