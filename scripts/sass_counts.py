"""SASS evidence per kernel family: instruction counts of the Blackwell-native paths (tcgen05 = UTC*MMA, TMA =
UTMALDG / UTMASTG / UBLKCP, TMEM loads/stores = LDTM / STTM) and of the legacy ones (HMMA = mma.sync, LDGSTS = cp.async)
in the built library.  usage: python scripts/sass_counts.py > profiles/sass_r2.md   (runs here, no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uplift_upsample_3dhpe_b200", "libuu3d.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCMMA other", r"\bUTC[A-GI-Z]MMA"),
            ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAPF", r"\bUTMAPF|UTMACCTL"), ("UBLKCP", r"\bUBLKCP"),
            ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS"), ("MUFU", r"\bMUFU"),
            ("SYNCS", r"\bSYNCS")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fam = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            raw = m.group(1)
            dem = subprocess.run(["c++filt", raw], capture_output=True, text=True).stdout.strip() or raw
            dem = re.sub(r"^void ", "", dem)
            name = re.sub(r"\(.*", "", dem)
            fam.setdefault(name, collections.Counter())["functions"] += 0
            fam[name]["_inst"] += 0
            continue
        if name is None or "/*" not in line:
            continue
        mm = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not mm:
            continue
        fam[name]["_inst"] += 1
        for key, pat in PATTERNS:
            if re.search(pat, mm.group(1)):
                fam[name][key] += 1
    # group template instantiations of one kernel
    groups = collections.OrderedDict()
    for n, c in fam.items():
        base = re.sub(r"<.*", "", n)
        g = groups.setdefault(base, collections.Counter())
        g.update(c)
        g["instantiations"] += 1
    keys = [k for k, _ in PATTERNS]
    print(f"source: `cuobjdump -sass {os.path.relpath(LIB, ROOT)}` (sm_100a), counts summed over the template instantiations of a kernel\n")
    print("| kernel | inst. | SASS instr | " + " | ".join(keys) + " |")
    print("|---|---:|---:|" + "---:|" * len(keys))
    for base, c in sorted(groups.items(), key=lambda kv: -sum(kv[1][k] for k in keys[:9])):
        if not any(c[k] for k in keys) and c["_inst"] < 200:
            continue
        print(f"| `{base}` | {c['instantiations']} | {c['_inst']} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")


if __name__ == "__main__":
    sys.exit(main())
