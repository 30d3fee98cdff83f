#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(out.splitlines())
    hdr = next(rd)
    units = next(rd)
    for row in rd:
        print("==", row[hdr.index("Kernel Name")][:70])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"   {h:64s} {row[i]:>16s} {units[i]}")
        for i, h in enumerate(hdr):
            if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    if float(row[i]) > 0.5:
                        print(f"   stall {h.split('warps_issue_stalled_')[1].split('_per_issue')[0]:40s} {float(row[i]):8.2f}")
                except ValueError:
                    pass


if __name__ == "__main__":
    main(sys.argv[1])
