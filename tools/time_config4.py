"""Timing of BASELINE config 4 (icp_rust_b200/workloads.py: pairs (frame i, frame i+1 mod 20) of the
bundled 3-D recording x 16,384 points, seeded initial perturbation, 20 iterations) on the streaming
kernel.  usage: time_config4.py [pairs=296] [check=16]   (env ICP_B200_STREAM_CELLS_PER_POINT sweeps the grid)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import icp_rust_b200 as icp
from icp_rust_b200.workloads import config4_pairs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
check = int(sys.argv[2]) if len(sys.argv) > 2 else 16
src, dst, off, T0 = config4_pairs(B)
ctx = icp.Context(0)
best = None
for rep in range(3):
    t = time.time()
    g = ctx.estimate_batch(src, off, dst, off, T0, 20, dim=3)
    dt = time.time() - t
    info = ctx.call_info()
    best = info["kernel_ms"] if best is None else min(best, info["kernel_ms"])
out = {"pairs": B, "cells_per_point": os.environ.get("ICP_B200_STREAM_CELLS_PER_POINT", "default"),
       "kernel_ms": best, "pairs_per_s": B / (best * 1e-3), "nn_queries_per_s": B * 16384 * 20 / (best * 1e-3),
       "grid": [info["grid_ctas"], info["block_threads"]], "accepted_per_outer_iteration": float((g["inner"] & 0xffff).mean()),
       "status_ok": bool((g["status"] == 0).all()), "redo_pairs": int(g["redo_pairs"])}
if check:
    import oracle
    n = min(B, check)
    o = oracle.icp_estimate_batch(3, src[:n * 16384], off[:n + 1], dst[:n * 16384], off[:n + 1], T0[:n], 20, nn_mode=oracle.NN_KDTREE)
    out["oracle_check"] = {"pairs": n, "inner_equal": bool(np.array_equal(o["inner"], g["inner"][:n])),
                           "corr_hash_equal": bool(np.array_equal(o["corr_hash"], g["corr_hash"][:n])),
                           "max_abs_transform_difference": float(np.abs(o["T"] - g["T"][:n]).max())}
print(json.dumps(out))
