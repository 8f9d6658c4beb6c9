#!/usr/bin/env python3
"""Per-source-line summary of an ncu report's source page (cuda,sass view): warp instructions executed,
avg active threads, stall samples. Usage: ncu_lines.py report.ncu-rep [min_share_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
items, fname, hdr = [], None, None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": fname = row[1].split("/")[-1]; hdr = None; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed"); continue
    if hdr is None or row[0] == "": continue     # SASS rows have an empty line number
    try: items.append((fname, row[0], int(row[ci] or 0), int(row[cs] or 0), int(row[ct] or 0), row[1]))
    except ValueError: pass
tot_i = sum(i[2] for i in items); tot_s = sum(i[3] for i in items)
print(f"total warp instructions {tot_i}, samples {tot_s}")
for f, ln, i, s, t, src in items:
    if i * 100.0 >= thresh * tot_i or s * 100.0 >= thresh * tot_s:
        print(f"{f}:{ln:>4} inst {100.0*i/tot_i:5.1f}% samp {100.0*s/max(tot_s,1):5.1f}% thr {t/max(i,1):5.1f} | {src.strip()[:110]}")
