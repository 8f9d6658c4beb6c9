#!/usr/bin/env python3
"""Attribute the local-memory instructions (LDL / STL) and the instruction count of one kernel to source lines.
usage: tools/sass_local.py <cubin> <kernel-name-substring>      (cubin: cuobjdump -xelf all <object>; built with -lineinfo)"""
import re
import subprocess
import sys
from collections import Counter

text = subprocess.run(["nvdisasm", "-g", sys.argv[1]], capture_output=True, text=True).stdout.splitlines()
inside, cur = False, None
local, total, per_file = Counter(), 0, Counter()
for l in text:
    if l.startswith("//---------------------"):
        inside = sys.argv[2] in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", l):
        total += 1
        per_file[cur[0] if cur else "?"] += 1
        if re.search(r"\b(LDL|STL)", l):
            local[(cur, "indexed" if "[R1" not in l else "spill")] += 1
print("instructions", total, dict(per_file))
for k, v in sorted(local.items(), key=lambda kv: (kv[0][0] or ("", 0))):
    print(k, v)
