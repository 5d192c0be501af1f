"""Timing of BASELINE config 3: one shared 100k-point 2-D map, N scans x 1080 points from random
poses with perturbed initial transforms, 20 outer iterations, each scan against the WHOLE map
(icp_b200_target_estimate_batch).  usage: time_config3.py [scans=65536] [check=64]
Prints one JSON line.  `check` scans are compared with the oracle (kd-tree) on the host."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import icp_rust_b200 as icp


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    check = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    m = icp.synth_map(100_000, seed=7)
    scans, off, T0 = icp.synth_map_scans(B, beams=1080, seed=7)
    ctx = icp.Context(0)
    t0 = time.perf_counter()
    tgt = icp.Target(m, ctx)
    build_ms = (time.perf_counter() - t0) * 1e3
    rows = []
    g = None
    for rep in range(3):
        t = time.perf_counter()
        g = tgt.estimate_batch(scans, off, T0, 20, want_inner=True, want_hash=True)
        wall = time.perf_counter() - t
        info = ctx.call_info()
        rows.append((wall * 1e3, info["kernel_ms"]))
    wall_ms, kernel_ms = min(r[0] for r in rows), min(r[1] for r in rows)
    out = {"workload": f"config3: {B} scans x 1080 points vs one shared 100000-point map, 20 outer iterations",
           "target_build_ms": round(build_ms, 2),
           "kernel_ms": round(kernel_ms, 2), "host_call_ms": round(wall_ms, 2),
           "pairs_per_s_device": round(B / (kernel_ms * 1e-3), 1), "pairs_per_s_host_call": round(B / (wall_ms * 1e-3), 1),
           "nn_queries_per_s_inside_the_loop": round(B * 1080 * 20 / (kernel_ms * 1e-3), 1),
           "gn_evaluations_per_pair": float((g["inner"] & 0xffff).sum(axis=1).mean()) if "inner" in g else None,
           "grid": [info["grid_ctas"], info["block_threads"], info["smem_bytes"]],
           "status_ok": bool((g["status"] == 0).all())}
    if check:
        import oracle
        worst = 0.0
        same_hash = True
        for b in range(check):
            o = oracle.icp_estimate(scans[off[b]:off[b + 1]], m, T0[b], 20, nn_mode=oracle.NN_KDTREE)
            worst = max(worst, float(np.abs(o["T"] - g["T"][b]).max()))
            same_hash &= int(o["corr_hash"]) == int(g["corr_hash"][b])
        out["oracle_check"] = {"scans": check, "max_abs_transform_difference": worst, "correspondence_hashes_equal": bool(same_hash)}
    tgt.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
