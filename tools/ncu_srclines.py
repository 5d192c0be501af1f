"""Per-source-line totals of an `ncu -i rep --page source --csv --print-source cuda,sass` dump:
share of warp instructions, average active lanes per instruction, share of stall samples.
usage: ncu_srclines.py <dump.csv> [kernel ordinal, default 0] [top N, default 40] [csrc dir: also per-region totals]
A source line belongs to the last `// @region <name>` marker above it in its file."""
import csv
import os
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    agg = defaultdict(lambda: [0, 0, 0])
    src = {}
    seen = defaultdict(int)
    take, hdr, cur = False, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            take = seen[cur] == which
            seen[cur] += 1
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if r[0] == "Function Name" or hdr is None or not take or r[2] != "-":
            continue
        try:
            ln = int(r[0])
            ie, te, smp = int(r[hdr["Instructions Executed"]]), int(r[hdr["Thread Instructions Executed"]]), int(r[hdr["# Samples"]])
        except ValueError:
            continue
        if ie == 0:
            continue
        a = agg[(cur, ln)]
        a[0] += ie; a[1] += te; a[2] += smp
        src[(cur, ln)] = r[1]
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[2] for v in agg.values()) or 1
    print(f"warp instructions {tot}, thread instructions {sum(v[1] for v in agg.values())}, samples {tots}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{k[0]:20s}{k[1]:4d} inst {100 * v[0] / tot:5.1f}%  lanes {v[1] / v[0]:5.1f}  samples {100 * v[2] / tots:5.1f}%  {src[k].strip()[:88]}")
    if len(sys.argv) > 4:
        csrc = sys.argv[4]
        reg = defaultdict(lambda: [0, 0, 0])
        rmap = {}
        for fn in os.listdir(csrc):
            if fn.endswith((".cuh", ".cu", ".h")):
                cur, m = fn, {}
                for ln, text in enumerate(open(os.path.join(csrc, fn), errors="replace"), 1):
                    g = re.search(r"//\s*@region\s+(\S+)", text)
                    if g:
                        cur = g.group(1)
                    m[ln] = cur
                rmap[fn] = m
        for (fn, ln), v in agg.items():
            r = rmap.get(fn, {}).get(ln, fn)
            for q in range(3):
                reg[r][q] += v[q]
        print("\nper region:")
        for r, v in sorted(reg.items(), key=lambda kv: -kv[1][0]):
            print(f"{r:28s} inst {100 * v[0] / tot:5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / tots:5.1f}%")


if __name__ == "__main__":
    main()
