"""Source-level stall summary of one kernel in an `ncu --set full --import-source on` capture (runs here, no GPU):
warp-stall reasons, the opcode mix with its share of the stall samples, and the hottest SASS instructions.
usage: python scripts/ncu_stalls.py gpurun_out/x.ncu-rep [title] > profiles/NAME.md"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
n = lambda r, k: int(r[col[k]] or 0)
tot = sum(n(r, "# Samples") for r in data)
print(f"# {title}\n\nkernel: `{kernel[:110]}`\nsource: `{rep}` (ncu --set full --clock-control none --import-source on; {tot} warp samples)\n")
rr = list(csv.reader(raw.splitlines()))
if len(rr) >= 3:
    h, v = rr[0], rr[-1]
    want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
    print("| metric | value | unit |\n|---|---:|---|")
    for w in want:
        if w in h:
            print(f"| `{w}` | {v[h.index(w)]} | {rr[1][h.index(w)]} |")
    print()
agg = {}
for r in data:
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            agg[k[6:]] = agg.get(k[6:], 0) + n(r, k)
print("Warp-stall reasons (share of all samples): " + ", ".join(f"{k} {100 * v / max(tot, 1):.1f} %" for k, v in
      sorted(agg.items(), key=lambda kv: -kv[1])[:9]) + "\n")
op = {}
for r in data:
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", r[col["Source"]])
    o = m.group(1) if m else "?"
    e = op.setdefault(o, [0, 0])
    e[0] += n(r, "# Samples"); e[1] += n(r, "Instructions Executed")
te = sum(v[1] for v in op.values())
print("| opcode | share of executed instructions | share of stall samples |\n|---|---:|---:|")
for k, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"| {k} | {100 * v[1] / max(te, 1):.1f} % | {100 * v[0] / max(tot, 1):.1f} % |")
print("\n| samples | executed | instruction | main stall |\n|---:|---:|---|---|")
for r in sorted(data, key=lambda r: -n(r, "# Samples"))[:14]:
    st = {k[6:]: n(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    top = max(st.items(), key=lambda kv: kv[1])
    print(f"| {n(r, '# Samples')} | {n(r, 'Instructions Executed')} | `{r[col['Source']].strip()[:80]}` | {top[0]} {top[1]} |")
