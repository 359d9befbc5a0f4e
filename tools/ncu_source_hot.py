"""Top stall-sample instructions of one kernel from an ncu report's source page (run here, no GPU needed):
   ncu -i rep.ncu-rep --page source --csv --kernel-name K | python tools/ncu_source_hot.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: k for k, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ci[s]] or 0) for r in data) for s in stalls}
print("total samples", tot)
print("stall mix:", {k: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
data.sort(key=lambda r: -int(r[ci["# Samples"]] or 0))
for r in data[:n]:
    s = int(r[ci["# Samples"]] or 0)
    top = max(stalls, key=lambda k: int(r[ci[k]] or 0))
    print(f"{s/tot:6.3f} {top:22s} {r[ci['Source']].strip()[:100]}")
