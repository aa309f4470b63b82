#!/usr/bin/env python
"""Condense an `ncu --set full` report (.ncu-rep) into a small per-launch table for profiles/.

    python tools/summarize_ncu.py gpurun_out/frame_c2.ncu-rep > profiles/r1_frame_c2_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_%"),
    ("sm__inst_executed_pipe_xu.sum", "xu_inst"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# source: %s (ncu --set full --clock-control none; per-launch, cold-cache, serialised)" % rep)
    for d in data:
        name = d[idx["Kernel Name"]]
        print("\n%s" % name[:140])
        for k, label in KEYS:
            if k in idx:
                print("  %-16s %18s %s" % (label, d[idx[k]], units[idx[k]]))


if __name__ == "__main__":
    main()
