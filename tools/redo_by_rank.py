"""Which of bench.py's per-rank input buffers (rank r, buffer j: pairs [(2r + j) * 4096, +4096)) hold a pair the
knife-edge guard re-runs, and what a host-pointer call on such a buffer costs beside one without."""
import json
import time

import numpy as np

import icp_rust_b200 as icp

ctx = icp.Context(0)
B = 4096
T0 = np.tile(np.array([1., 0., 0., 1., 0., 0.]), (B, 1))
rows = []
for r in range(8):
    for j in range(2):
        first = (r * 2 + j) * B
        src, dst, off, _ = icp.synth_scan_pairs(B, beams=1080, first_pair=first)
        ctx.estimate_batch(src, off, dst, off, T0, 20, want_inner=False, want_hash=False)
        t0 = time.perf_counter()
        g = ctx.estimate_batch(src, off, dst, off, T0, 20, want_inner=False, want_hash=False)
        ms = (time.perf_counter() - t0) * 1e3
        rows.append({"rank": r, "buffer": j, "redo_pairs": int(g["redo_pairs"]), "call_ms_pageable": round(ms, 2),
                     "kernel_ms": round(ctx.call_info()["kernel_ms"], 2)})
        print(json.dumps(rows[-1]), flush=True)
