"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (run here, no GPU needed)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        val = val / 1e3 if unit == "ns" else (val * 1e3 if unit == "ms" else val)
        agg[name].append(val)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'avg_us':>10s} {'total_ms':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:60]:60s} {len(v):8d} {sum(v)/len(v):10.1f} {sum(v)/1e3:9.2f} {sum(v)/tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
