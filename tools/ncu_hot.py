"""Hot source lines of a kernel from an ncu report captured with --import-source on (-lineinfo build).
usage: ncu_hot.py <rep.ncu-rep> [top=40]
Per CUDA source line: warp instructions executed (share), average active threads, stall samples (share) and
the dominant stall reasons of those samples."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), ""])
cur = None; col = None
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; col = {h: i for i, h in enumerate(hdr)}; stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; continue
    if col is None or not r[0].strip().isdigit(): continue
    try:
        k = (cur, int(r[0]))
        a = agg[k]
        a[0] += int(r[col["Instructions Executed"]] or 0)
        a[1] += int(r[col["Thread Instructions Executed"]] or 0)
        a[2] += int(r[col["# Samples"]] or 0)
        for h in stalls:
            v = r[col[h]]
            if v and v != "0": a[3][h] += int(v)
        a[4] = r[1]
    except (ValueError, IndexError):
        pass
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"total warp instructions {ti:,}  samples {ts:,}")
print("by stall samples:")
for k, a in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    st = ", ".join(f"{h[6:]} {v * 100 // max(a[2], 1)}%" for h, v in a[3].most_common(3))
    print(f"{k[0]}:{k[1]:<5} inst {a[0] * 100 / max(ti, 1):5.1f}%  lanes {a[1] / max(a[0], 1):5.1f}  samples {a[2] * 100 / max(ts, 1):5.1f}%  [{st}]  {a[4].strip()[:70]}")
