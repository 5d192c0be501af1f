"""Which 512-pair chunk of bench.py's rank-r buffer 1 holds the pair the guard re-runs."""
import numpy as np
import icp_rust_b200 as icp
ctx = icp.Context(0)
B = 4096
for r in (2, 5, 6):
    first = (r * 2 + 1) * B
    src, dst, off, _ = icp.synth_scan_pairs(B, beams=1080, first_pair=first)
    for c in range(8):
        s = src[c * 512 * 1080:(c + 1) * 512 * 1080]; d = dst[c * 512 * 1080:(c + 1) * 512 * 1080]
        o = (off[c * 512:(c + 1) * 512 + 1] - off[c * 512]).astype(np.uint32)
        g = ctx.estimate_batch(s, o, d, o, np.tile(np.array([1., 0, 0, 1, 0, 0]), (512, 1)), 20, want_inner=False, want_hash=False)
        if g["redo_pairs"]:
            print("rank", r, "chunk", c, "redo", g["redo_pairs"], flush=True)
