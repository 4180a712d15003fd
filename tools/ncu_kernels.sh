#!/bin/bash
# One ncu --set full capture of the hot kernels at 8192^2 (run after the same command exited 0 without ncu).
# usage: tools/ncu_kernels.sh <tag> ; output gpurun_out/prof_<tag>.ncu-rep + raw csv of selected metrics
tag=$1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${tag}_plain.json 2> gpurun_out/ncu_${tag}_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_down5|k_up5|k_smooth9|k_spmv5|k_update|k_residual9|k_direction" -s 40 -c 12 \
    -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
python - <<P
import csv
rows = list(csv.reader(open("gpurun_out/prof_${tag}_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(w) for w in want if w in hdr]
print([hdr[i] for i in idx])
for r in rows[2:]:
    print([r[i][:60] for i in idx])
P
