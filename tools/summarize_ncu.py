"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    us = v / 1000 if r["Metric Unit"] in ("ns", "nsecond") else v
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us serialized\n")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0]:.1f} |")
