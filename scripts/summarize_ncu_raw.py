"""Summarise an `ncu --page raw --csv` export of one forward pass: per-kernel table (markdown) and the DRAM
traffic per launch by kernel kind (JSON, read by bench.py for roofline.traffic).
usage: python scripts/summarize_ncu_raw.py gpurun_out/prof_raw.csv BATCH profiles/NAME [LAST_N] [TRAFFIC_JSON]  (writes
NAME.md and TRAFFIC_JSON (default ncu_traffic_r2.json) next to it; LAST_N keeps only the last N launches = the final
forward pass of a longer capture)"""
import collections
import csv
import json
import os
import re
import sys

path, batch, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
if "Metric Name" in rows[0]:          # long format of `ncu --csv --log-file` (one row per launch and metric) -> wide
    h0 = rows[0]
    iid, ik, imn, imu, imv = (h0.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    names, unit_of, by_id = [], {}, collections.OrderedDict()
    for r in rows[1:]:
        if r[imn] not in unit_of:
            names.append(r[imn]); unit_of[r[imn]] = r[imu]
        by_id.setdefault(r[iid], {"Kernel Name": r[ik]})[r[imn]] = r[imv]
    rows = [["Kernel Name"] + names, [""] + [unit_of[n] for n in names]] + \
           [[d["Kernel Name"]] + [d.get(n, "") for n in names] for d in by_id.values()]
hdr, units, data = rows[0], rows[1], rows[2:]
if len(sys.argv) > 4 and int(sys.argv[4]) > 0:
    data = data[-int(sys.argv[4]):]
traffic_name = sys.argv[5] if len(sys.argv) > 5 else "ncu_traffic_r2.json"
col = {h: i for i, h in enumerate(hdr)}


def scale(v, u, kind):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else 0.0
    if kind == "time":
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)            # -> us
    if kind == "bytes":
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)  # -> bytes
    return v


def get(r, name, kind=None):
    if name not in col:
        return 0.0
    return scale(r[col[name]], units[col[name]], kind)


KIND = [("k_gemm_tc", "gemm_tc"), ("k_mlp_tc", "gemm_tc"), ("k_spatial_tc", "spatial"), ("k_attention_tc", "attention"),
        ("k_residual_ln", "layernorm"), ("k_layernorm", "layernorm"), ("k_token_fill", "token_fill"),
        ("k_mask", "gather"), ("k_window", "gather")]
agg = collections.OrderedDict()
kinds = collections.OrderedDict()
for r in data:
    full = r[col["Kernel Name"]]
    name = re.sub(r"\(.*", "", full).replace("void ", "").replace("uu::", "")
    t = get(r, "gpu__time_duration.sum", "time")
    rd, wr = get(r, "dram__bytes_read.sum", "bytes"), get(r, "dram__bytes_write.sum", "bytes")
    if rd + wr == 0 and "dram__bytes.sum.per_second" in col:      # section sets report the rate: bytes = rate x duration
        u = units[col["dram__bytes.sum.per_second"]]
        rate = get(r, "dram__bytes.sum.per_second") * {"byte/s": 1.0, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9,
                                                       "Tbyte/s": 1e12}.get(u, 1.0)
        rd, wr = rate * t * 1e-6, 0.0
    a = agg.setdefault(name, collections.defaultdict(float))
    a["n"] += 1; a["us"] += t; a["rd"] += rd; a["wr"] += wr
    for k, m in (("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 ("dram", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("l2", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("regs", "launch__registers_per_thread")):
        a[k] += get(r, m) * t                     # time-weighted
    kind = next((v for k, v in KIND if k in name), "other")
    kk = kinds.setdefault(kind, collections.defaultdict(float))
    kk["n"] += 1; kk["us"] += t; kk["bytes"] += rd + wr
tot = sum(a["us"] for a in agg.values())
with open(out + ".md", "w") as f:
    f.write(f"source: {path} ({len(data)} launches = one forward pass, B = {batch} windows; ncu --metrics gpu__time_duration, "
            "sm__pipe_tensor_cycles_active, smsp__issue_active, gpu__dram_throughput, lts__throughput, dram__bytes_read / "
            "write, registers; --clock-control none; per-launch times are cold-cache and serialised: compare shares)\n\n")
    f.write("| kernel | launches | total us | share | DRAM MB (read+write, or read if split) | DRAM write MB | DRAM GB/s | tensor pipe % | issue % | DRAM % | L2 % | regs |\n")
    f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        w = a["us"] or 1.0
        f.write(f"| `{k}` | {int(a['n'])} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | "
                f"{(a['rd'] + a['wr']) / a['us'] / 1e3:.0f} | {a['tensor'] / w:.1f} | {a['issue'] / w:.1f} | {a['dram'] / w:.1f} | "
                f"{a['l2'] / w:.1f} | {a['regs'] / w:.0f} |\n")
    f.write(f"\ntotal: {tot:.1f} us\n")
json.dump({"source": os.path.basename(path), "batch": batch,
           "kinds": {k: {"launches": int(v["n"]), "us": round(v["us"], 1),
                         "dram_bytes_per_launch": round(v["bytes"] / v["n"])} for k, v in kinds.items()}},
          open(os.path.join(os.path.dirname(out), traffic_name), "w"), indent=1)
print(open(out + ".md").read())
