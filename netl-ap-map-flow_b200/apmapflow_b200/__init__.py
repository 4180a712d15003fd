"""apmapflow_b200 -- B200-native drop-in for the LCL hot path of stadelmanma/netl-ap-map-flow.

Mirrors the reference's ``apmapflow.run_model`` surface (InputFile, run_model, BulkRun,
estimate_req_RAM; apmapflow/run_model/__init__.py:18-20) on top of libapmlcl.so.
"""
__version__ = "0.1.0"
