#!/usr/bin/env python
"""Key metrics of every launch in an ncu report:  python tools/ncu_launch_metrics.py <report.ncu-rep>"""
import csv, io, subprocess, sys
M = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = rows[0]
print("%-34s %9s %12s %9s %9s %7s %7s %5s %6s" % ("kernel", "us", "warp inst", "rd MB", "wr MB", "issue%", "warps%", "regs", "dram%"))
for r in rows[2:]:
    d = dict(zip(H, r))
    print("%-34s %9.1f %12.0f %9.1f %9.1f %7.1f %7.1f %5s %6.1f" % (d["Kernel Name"][:34], float(d[M[0]]), float(d[M[1]]), float(d[M[2]]), float(d[M[3]]),
          float(d[M[4]]), float(d[M[5]]), d[M[6]], float(d[M[7]])))
