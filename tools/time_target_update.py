"""icp_b200_target_update against a full rebuild on the config-3 map (100,000 points): 1 % of the points
replaced per update, inside the bounding box.  Prints one JSON line (wall-clock per call, host included)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import icp_rust_b200 as icp

ctx = icp.Context(0)
m = icp.synth_map(100_000, seed=7)
rng = np.random.default_rng(1)
t = icp.Target(m, ctx)
pts = m.copy()
scans, off, T0 = icp.synth_map_scans(64, beams=1080, seed=7)
q = np.concatenate([scans[off[b]:off[b + 1]] @ T0[b, :4].reshape(2, 2).T + T0[b, 4:6] for b in range(64)])
res = {}
for mode in ("1", "0"):
    os.environ["ICP_B200_TARGET_INCREMENTAL"] = mode
    ts, ss = [], []
    for rep in range(12):
        rem = rng.choice(len(pts), 1000, replace=False).astype(np.uint32)
        add = np.clip(pts[rng.integers(0, len(pts), 1000)] + rng.normal(0, 10, (1000, 2)), m.min(0), m.max(0))
        t0 = time.perf_counter(); t.update(add, rem); ts.append(time.perf_counter() - t0)
        keep = np.ones(len(pts), bool); keep[rem] = False
        pts = np.concatenate([pts[keep], add])
        t0 = time.perf_counter(); got = t.search(q); ss.append(ctx.call_info()["kernel_ms"])
    res["incremental" if mode == "1" else "rebuild"] = {"update_ms_median": float(np.median(ts[2:]) * 1e3),
                                                         "search_kernel_ms_after": float(np.median(ss[2:]))}
fresh = icp.Target(pts, ctx)
res["indices_equal_fresh_build"] = bool(np.array_equal(t.search(q), fresh.search(q)))
t0 = time.perf_counter(); f2 = icp.Target(pts, ctx); res["target_create_ms"] = (time.perf_counter() - t0) * 1e3
res["points"] = len(pts); res["queries"] = len(q)
print(json.dumps(res))
