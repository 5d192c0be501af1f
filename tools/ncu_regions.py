"""Per-region totals of an ncu SASS-level source dump.  A source line belongs to the last
`// @region <name>` marker above it in its file (files without markers form one region each).

usage: ncu_regions.py <ncu_source.csv> <nvdisasm -g .sass> <kernel name substring> [csrc dir]
"""
import csv
import os
import re
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines


def region_map(csrc):
    out = {}
    for fn in os.listdir(csrc):
        if not fn.endswith((".cuh", ".cu", ".h")):
            continue
        cur = fn
        m = {}
        for ln, text in enumerate(open(os.path.join(csrc, fn), errors="replace"), 1):
            g = re.search(r"//\s*@region\s+(\S+)", text)
            if g:
                cur = g.group(1)
            m[ln] = cur
        out[fn] = m
    return out


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    csrc = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "icp_rust_b200", "csrc")
    regions = region_map(csrc)
    a2l = ncu_lines.load_lines(sass, kernel)
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = defaultdict(lambda: defaultdict(float))
    base = None
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[col["Address"]], 16)
        if base is None:
            base = addr
        (f, l), _ = a2l.get(addr - base, (("?", 0), ""))
        name = regions.get(f, {}).get(l, f)
        d = agg[name]
        d["samples"] += float(r[col["# Samples"]] or 0)
        d["inst"] += float(r[col["Instructions Executed"]] or 0)
        for c in stall_cols:
            d[c] += float(r[col[c]] or 0)
    ts = sum(d["samples"] for d in agg.values())
    ti = sum(d["inst"] for d in agg.values())
    print(f"total samples {ts:.0f}  warp instructions {ti:.4e}")
    print(f"{'region':28s} {'samp%':>6s} {'inst%':>6s}  top stalls (share of region samples)")
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        st = sorted(((d[c], c[6:]) for c in stall_cols), reverse=True)[:4]
        sts = " ".join(f"{n}:{v / max(d['samples'], 1) * 100:.0f}%" for v, n in st if v > 0)
        print(f"{name:28s} {d['samples'] / ts * 100:6.2f} {d['inst'] / ti * 100:6.2f}  {sts}")


if __name__ == "__main__":
    main()
