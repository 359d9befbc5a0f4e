"""Aggregate warp-stall samples per CUDA source line:
   ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv | python tools/ncu_source_lines.py [N]"""
import csv
import sys
from collections import defaultdict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = list(csv.reader(sys.stdin))
agg = defaultdict(int)
text = {}
cur_file = ""
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        li, si, ni = 0, 1, hdr.index("# Samples")
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    try:
        line = int(r[li])
        s = int(r[ni] or 0)
    except ValueError:
        continue
    if r[2] == "":            # the CUDA source row itself (SASS rows carry an address)
        text[(cur_file, line)] = r[si].strip()
        agg[(cur_file, line)] += 0
    else:
        agg[(cur_file, line)] += s
tot = sum(agg.values()) or 1
print("total samples", tot)
for (f, line), s in sorted(agg.items(), key=lambda kv: -kv[1])[:n]:
    print(f"{s / tot:6.3f}  {f}:{line:<5d} {text.get((f, line), '')[:110]}")
