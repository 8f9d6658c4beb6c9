#!/usr/bin/env python3
"""Per-instruction listing of one kernel of an .ncu-rep with execution counts, and a summary by execution-count level.
usage: tools/ncu_segments.py report.ncu-rep kernel_regex [listing_out.txt]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ie, it, isamp, isrc = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples", "Source"))
seen, data = set(), []
for r in rows[2:]:
    if len(r) > 20 and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); data.append((int(r[ia], 16), int(r[ie]), int(r[it]), int(r[isamp]), r[isrc]))
base = data[0][0]
tot = sum(r[1] for r in data); tots = sum(r[3] for r in data)
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        for a, e, t, s, src in data:
            f.write(f"{a - base:6x} {e / 1e6:8.3f}M {t / max(e, 1):5.1f} {s:6d}  {src}\n")
print(f"total warp instructions {tot}, SASS instructions {len(data)}")
seg, cur = [], None
for a, e, t, s, src in data:
    if cur is None or abs(e - cur[2]) > 0.12 * max(cur[2], 0.0005 * tot / 100):
        cur = [a - base, a - base, e, 0, 0, 0, 0]; seg.append(cur)
    cur[1] = a - base; cur[3] += e; cur[4] += 1; cur[5] += t; cur[6] += s
for s in seg:
    if s[3] / tot > 0.004:
        print(f"{s[0]:6x}-{s[1]:6x} exec {s[2] / 1e6:8.3f}M n={s[4]:4d} inst share {s[3] / tot:6.3f} thr/inst {s[5] / max(s[3], 1):5.1f} samples {s[6] / max(tots, 1):6.3f}")
