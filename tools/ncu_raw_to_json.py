"""Turns the raw page of an ncu capture of tools/ncu_roofline.py (`ncu -i x.ncu-rep --page raw --csv | gzip`) into the
per-kernel summary bench.py and profiles/README.md use (profiles/r02_ncu_traffic*.json).

    python tools/ncu_raw_to_json.py gpurun_out/prof_r02_final_raw.csv.gz 8192 8192 > profiles/r02_ncu_traffic_final.json
"""
import csv
import gzip
import json
import re
import sys

NAMES = ["spmv", "update", "direction", "down5", "up5", "smooth9_l1", "residual9_l1", "restrict1", "smooth9_l1_up"]   # launch order
COLS = {
    "grid": "launch__grid_size",
    "duration_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "inst_executed": "smsp__inst_executed.sum",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "registers": "launch__registers_per_thread",
    "shared_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "shared_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "stall_long_sb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short_sb": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_mio": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_math_pipe": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    path, size, nrows = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rows = list(csv.reader(gzip.open(path, "rt") if path.endswith(".gz") else open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    assert len(data) == len(NAMES), (len(data), "launches captured, expected", len(NAMES))
    kernels = {}
    for name, r in zip(NAMES, data):
        k = {"ncu_kernel": re.sub(r"^void ", "", r[col["Kernel Name"]]).split("(")[0]}
        for key, metric in COLS.items():
            i = col[metric]
            v = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            k[key] = int(round(v)) if key in ("grid", "registers", "inst_executed", "shared_wavefronts", "shared_bank_conflicts",
                                               "dram_read_bytes", "dram_write_bytes") else round(v, 2)
        k["dram_GBps"] = round((k["dram_read_bytes"] + k["dram_write_bytes"]) / k["duration_us"] / 1e3, 1)
        kernels[name] = k
    out = {"_comment": "per launch, from one `APM_LCL_TIME_WARMUP=0 ncu --set full --clock-control none --nvtx --nvtx-include "
                       "roofline/ python tools/ncu_roofline.py %d %d` capture on B200 after the same command had exited 0 without "
                       "ncu; each kernel launched alone on the solved system of the bench map (%d rows x %d); converted by "
                       "tools/ncu_raw_to_json.py; bench.py copies dram_read_bytes + dram_write_bytes of the dominant kernel into "
                       "roofline.traffic" % (size, nrows, nrows, size),
           "size": size, "rows": nrows, "kernels": kernels}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
