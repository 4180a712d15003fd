#!/bin/bash
# End-of-round measurement pass (run on the GPU box through gpurun); everything lands in gpurun_out/.
O=gpurun_out
python bench.py > $O/bench_r01e.json 2> $O/bench_r01e.err
python bench.py --impl reference > $O/bench_ref_r01e.json 2>> $O/bench_r01e.err
for n in 1024 2048 4096 8192; do for c in 0 0.1 0.2 0.3; do
  python bench.py --size $n --contact $c --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$n', '$c', round(d['ms_per_step'],2), d['config']['pcg_iterations'], round(d['value'],1))"
done; done > $O/table_r01e.txt
python bench.py --size 16384 --contact 0.2 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('16384 0.2', round(d['ms_per_step'],2), d['config']['pcg_iterations'], round(d['value'],1), d['e2e'])" >> $O/table_r01e.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file $O/launches_r01e.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches_r01e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_down5|k_up5|k_smooth9|k_spmv5|k_update|k_residual9|k_tail_cycle" -s 60 -c 14 \
    -o $O/prof_r01e -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full_r01e.log 2>&1
tail -2 $O/ncu_full_r01e.log; cat $O/table_r01e.txt; cut -c1-300 $O/bench_r01e.json; cat $O/bench_ref_r01e.json | cut -c1-300
