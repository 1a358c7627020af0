#!/usr/bin/env python
"""Print selected metrics of an .ncu-rep (raw page) per kernel launch.  usage: ncu_pick.py file.ncu-rep [regex ...]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pats = [re.compile(p) for p in (sys.argv[2:] or [
    r"^gpu__time_duration.sum$", r"launch__(grid_size|block_size|registers_per_thread|occupancy_limit)", r"launch__waves",
    r"sm__warps_active.avg.pct_of_peak_sustained_active", r"smsp__issue_active.avg.pct", r"sm__inst_executed.sum$",
    r"smsp__inst_executed.sum$", r"sm__inst_executed_pipe_(fp64|fma|alu|xu|lsu|fmaheavy|uniform).*sum$",
    r"sm__pipe_fp64_cycles_active.avg.pct", r"sm__inst_executed_pipe_fp64.*pct", r"smsp__average_warp.*stall|smsp__average_warps_issue_stalled.*_per_issue_active",
    r"dram__bytes_(read|write).sum$", r"gpu__dram_throughput.avg.pct", r"sm__throughput.avg.pct", r"l1tex__t_sector_hit_rate", r"lts__t_sector_hit_rate.pct",
    r"smsp__warp_issue_stalled.*per_warp_active"])]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
for r in data:
    print("==", r[ki][:90])
    for i, h in enumerate(hdr):
        if any(p.search(h) for p in pats):
            print("   %-95s %s %s" % (h, r[i], units[i]))
