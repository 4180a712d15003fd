#!/bin/bash
# W-level knob sweep on one GPU: which mid levels need the second visit? usage: tools/w_sweep.sh <size> "<contacts>" "<lo:hi pairs>"
size=$1; contacts=$2; pairs=$3
for c in $contacts; do for p in $pairs; do
  lo=${p%%:*}; hi=${p##*:}
  APM_LCL_W_LO=$lo APM_LCL_W_HI=$hi python bench.py --size $size --contact $c --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']
print('size $size contact $c W $lo..$hi: ms', round(d['ms_per_step'],2), 'iters', d['config']['pcg_iterations'], 'vcycle ms', round(k['vcycle']['ms'],3))"
done; done
