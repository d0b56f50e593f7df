#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`): the counters DESIGN.md cites."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct" ]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print("kernels:", [r[ki].split("(")[0] for r in data])
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    for k in KEYS + stall:
        if k in hdr:
            i = hdr.index(k)
            print("%-95s %-10s %s" % (k, units[i], [r[i] for r in data]))
if __name__ == "__main__":
    main(sys.argv[1])
