"""Command-line callers of the LCL path, same names and arguments as apmapflow/scripts/
(``apm_run_lcl_model``, ``apm_bulk_run``); each module has a ``main()`` like the reference's console
entry points (setup.py:26-29 of the reference builds them as ``<name> = apmapflow.scripts.<name>:main``)."""
