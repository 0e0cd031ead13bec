#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into one row per captured launch with the metrics the
roofline argument needs.  Usage: python benchmarks/ncu_summary.py report.ncu-rep > profiles/x.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_selected"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    out = csv.writer(sys.stdout)
    out.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in idx])
    for r in rows[2:]:
        out.writerow([r[i][:90] for _, i in idx])


if __name__ == "__main__":
    main()
