"""Joins an ncu `--page source --csv` dump (SASS level, per-instruction stall samples) with
`nvdisasm -g` line info of the same kernel and prints per-source-line totals.

usage: ncu_lines.py <ncu_source.csv> <nvdisasm.sass> <mangled kernel name substring> [top N]
"""
import csv
import re
import sys
from collections import defaultdict


def load_lines(sass_path, kernel):
    addr2line = {}
    cur = None
    active = False
    for ln in open(sass_path, errors="replace"):
        if ln.startswith("//---") and ".text." in ln:
            active = kernel in ln
            cur = None
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return addr2line


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    a2l = load_lines(sass, kernel)
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    per_line = defaultdict(lambda: defaultdict(float))
    base = None
    tot_samples = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            addr = int(r[col["Address"]], 16)
        except ValueError:
            continue
        if base is None:
            base = addr
        off = addr - base
        key, _ = a2l.get(off, (("?", 0), ""))
        d = per_line[key]
        s = float(r[col["# Samples"]] or 0)
        d["samples"] += s
        tot_samples += s
        d["inst"] += float(r[col["Instructions Executed"]] or 0)
        d["conflicts"] += float(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        for c in stall_cols:
            d[c] += float(r[col[c]] or 0)
    tot_inst = sum(d["inst"] for d in per_line.values())
    print(f"total samples {tot_samples:.0f}, warp instructions {tot_inst:.3e}")
    print(f"{'file:line':32s} {'samp%':>6s} {'inst%':>6s} {'smem-excess':>12s}  top stalls")
    for key, d in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((d[c], c[6:]) for c in stall_cols), reverse=True)[:3]
        sts = " ".join(f"{n}:{v / max(d['samples'], 1) * 100:.0f}%" for v, n in st if v > 0)
        print(f"{key[0] + ':' + str(key[1]):32s} {d['samples'] / tot_samples * 100:6.2f} {d['inst'] / tot_inst * 100:6.2f} {d['conflicts']:12.3e}  {sts}")


if __name__ == "__main__":
    main()
