#!/bin/bash
# Round-2 measurement pass on ONE GPU (through gpurun); everything lands in gpurun_out/.
O=gpurun_out
python bench.py > $O/bench_r02.json 2> $O/bench_r02.err
python bench.py --impl reference > $O/bench_ref_r02.json 2>> $O/bench_r02.err
cut -c1-400 $O/bench_r02.json; cut -c1-300 $O/bench_ref_r02.json
# launch list of the same command (after it exited 0 above)
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_r02.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches_r02.log 2>&1
python tools/summarise_launches.py $O/launches_r02.csv | head -30
# full capture of each hot kernel, 8192^2 and the 2048 x 16384 strip of a 16384^2 / 8 solve
python tools/ncu_roofline.py 8192 > $O/ncu_roofline_plain.log 2>&1 && tail -10 $O/ncu_roofline_plain.log
# (the .ncu-rep files stay on the box: gpurun_out/ must stay under 64 MiB; the raw page travels as csv)
for cfg in "8192 8192" "16384 2048"; do set -- $cfg
  APM_LCL_TIME_WARMUP=0 ncu --set full --clock-control none --nvtx --nvtx-include "roofline/" -o /tmp/prof_r02_$1x$2 -f \
      python tools/ncu_roofline.py $1 $2 > $O/ncu_roofline_$1x$2.log 2>&1
  tail -3 $O/ncu_roofline_$1x$2.log
  ncu -i /tmp/prof_r02_$1x$2.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/prof_r02_$1x$2_raw.csv.gz
done
gzip -f $O/launches_r02.csv; rm -f $O/*.ncu-rep; du -sh $O
