#!/bin/bash
# Closing ncu pass of round 2 on ONE GPU (through gpurun), after bench.py exited 0 without ncu on the same code
# (tools/r02_final.sh): launch list of the bench command, then one full capture per hot kernel at 8192^2.
O=gpurun_out
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_r02_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches_r02_final.log 2>&1; echo "launch list rc=$?"
python tools/summarise_launches.py $O/launches_r02_final.csv | head -24
APM_LCL_TIME_WARMUP=0 timeout 300 ncu --set full --clock-control none --nvtx --nvtx-include "roofline/" -o /tmp/prof_r02_final -f \
    python tools/ncu_roofline.py 8192 8192 > $O/ncu_roofline_r02_final.log 2>&1; echo "full capture rc=$?"
tail -12 $O/ncu_roofline_r02_final.log
ncu -i /tmp/prof_r02_final.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/prof_r02_final_raw.csv.gz
gzip -f $O/launches_r02_final.csv; ls -la $O/*r02_final*
