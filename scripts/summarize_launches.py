"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step).
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [last_n_launches] > profiles/xxx.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
last_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
if last_n:
    rows = rows[-last_n:]
agg, tot = collections.OrderedDict(), 0.0
for x in rows:
    name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")
    v = float(x["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[x["Metric Unit"]]
    a = agg.setdefault(name, [0.0, 0, x["Grid Size"], x["Block Size"]])
    a[0] += v
    a[1] += 1
    tot += v
print(f"source: {path} ({len(rows)} launches, ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
print("| kernel | launches | total us | share | avg us | grid (last) | block |")
print("|---|---:|---:|---:|---:|---|---|")
for k, (v, n, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% | {v / n:.1f} | {g} | {b} |")
print(f"\ntotal: {tot:.1f} us")
