"""Latency of ONE reference call — Icp2d::new(dst).estimate(src, identity, 20) on the bundled scans 001 -> 002 and
on a 1080-point synthetic pair — through the single-pair entry of the C ABI (host buffers), per kernel instance
(ICP_B200_FORCE_INSTANCE) and reduction mode.  Prints JSON lines."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import icp_rust_b200 as icp
import oracle

z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scans2d.npz"))
pts, off = z["points"], z["offsets"]
cases = {"bundled 001->002": (pts[off[1]:off[2]].copy(), pts[off[2]:off[3]].copy())}
s, d, o_, _ = icp.synth_scan_pairs(1, beams=1080, first_pair=7)
cases["synthetic 1080"] = (s, d)
I = icp.Transform.identity()
for inst in (os.environ.get("INSTANCES", "-1,2,3,5")).split(","):
    if inst != "-1":
        os.environ["ICP_B200_FORCE_INSTANCE"] = inst
    else:
        os.environ.pop("ICP_B200_FORCE_INSTANCE", None)
    ctx = icp.Context(0)
    for name, (src, dst) in cases.items():
        ref = oracle.icp_estimate(src, dst, None, 20, nn_mode=oracle.NN_KDTREE)
        t = time.perf_counter()
        for _ in range(5):
            oracle.icp_estimate(src, dst, None, 20, nn_mode=oracle.NN_KDTREE)
        cpu_ms = (time.perf_counter() - t) / 5 * 1e3
        est = icp.Icp2d.new(dst, ctx)
        row = {"instance": inst, "pair": name, "cpu_oracle_ms": round(cpu_ms, 3)}
        for ordered in (False, True):
            est.estimate(src, I, 20, ordered=ordered)
            t = time.perf_counter()
            for _ in range(20):
                T = est.estimate(src, I, 20, ordered=ordered)
            row["ordered_ms" if ordered else "default_ms"] = round((time.perf_counter() - t) / 20 * 1e3, 3)
            row["ordered_bits_equal" if ordered else "default_max_abs_diff"] = (bool(np.array_equal(T.to_array(), ref["T"])) if ordered
                                                                               else float(np.abs(T.to_array() - ref["T"]).max()))
        row["block_threads"] = ctx.call_info()["block_threads"]
        print(json.dumps(row))
    ctx.close()
