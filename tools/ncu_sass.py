#!/usr/bin/env python3
"""SASS listing of an ncu report with per-instruction execution share, avg active threads and stall samples.
Usage: ncu_sass.py report.ncu-rep [min_share_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ci = hdr.index("Instructions Executed"); ct = hdr.index("Thread Instructions Executed"); cs = hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
tot = sum(int(r[ci]) for r in body)
print("total warp instructions", tot)
for n, r in enumerate(body):
    i = int(r[ci])
    if 100.0 * i / tot >= thresh:
        print(f"{n:5d} {100*i/tot:5.2f} {int(r[ct])/max(i,1):5.1f} {r[cs]:>5} {r[1][:100]}")
