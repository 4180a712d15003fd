#!/bin/bash
# A/B of two builds of libapmlcl.so on the same box: tools/ab_bench.sh <variant.so> <size> [bench args]
var=$1; size=$2; shift 2
for lib in "" "$var"; do
  APM_LCL_LIB=$lib python bench.py --size $size --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']
print('lib=${lib:-default} size $size: ms', round(d['ms_per_step'],3), 'iters', d['config']['pcg_iterations'], 'vcycle', round(k['vcycle']['ms'],3), {n: round(v['ms'],4) for n,v in k.items() if n!='vcycle'})"
done
