#!/bin/bash
# Round-2 closing pass on ONE GPU (through gpurun): GPU tests, the default bench line of the committed code, the
# reference arm, and BASELINE config 5 with and without shared unit solutions.  Everything lands in gpurun_out/.
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_final.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_r02_final.log
timeout 200 python bench.py > $O/bench_r02_final.json 2> $O/bench_r02_final.err; echo "bench rc=$?"; cut -c1-500 $O/bench_r02_final.json
timeout 120 python tools/gpu_bulk_bench.py 1024 16 > $O/bulk_r02_final.log 2>&1; echo "bulk rc=$?"; cat $O/bulk_r02_final.log
