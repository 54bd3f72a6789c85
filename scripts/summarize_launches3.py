"""Per-kernel table from an ncu launch list with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum.
usage: python scripts/summarize_launches3.py gpurun_out/launches.csv "title" > profiles/xxx.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else ""
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.OrderedDict()
for x in csv.DictReader(lines):
    d = per.setdefault(x["ID"], {"name": re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")})
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    if x["Metric Name"].startswith("gpu__time"):
        d["us"] = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[u]
    else:
        d[x["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
agg = collections.OrderedDict()
tot = 0.0
for d in per.values():
    a = agg.setdefault(d["name"], [0.0, 0, 0.0])
    a[0] += d.get("us", 0.0)
    a[1] += 1
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot += d.get("us", 0.0)
print(f"source: {path} — {title}: {len(per)} launches, {tot / 1e3:.2f} ms of kernel time (ncu --metrics gpu__time_duration.sum,"
      f"dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; cold-cache, serialised: compare shares).\n")
print("| kernel | launches | total us | share | avg us | DRAM MB / launch | DRAM GB/s |")
print("|---|---:|---:|---:|---:|---:|---:|")
for k, (v, n, b) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% | {v / n:.1f} | {b / n / 1e6:.1f} | {b / max(v, 1e-9) / 1e3:.0f} |")
