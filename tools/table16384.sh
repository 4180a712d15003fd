#!/bin/bash
# 16384^2 row of the size x contact table (gpurun_out/table16384.txt): ms per solve, iterations, Mcells/s
for c in 0 0.1 0.3; do
  python bench.py --size 16384 --contact $c --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('16384', '$c', round(d['ms_per_step'],2), d['config']['pcg_iterations'], round(d['value'],1))"
done | tee gpurun_out/table16384.txt
