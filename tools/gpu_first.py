"""First-light script for the GPU box: parity spot checks + a rough timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import icp_rust_b200 as icp

ctx = icp.Context(0)
z = np.load("tests/golden/scans2d.npz"); pts, off, ids = z["points"], z["offsets"], z["ids"]
sc = {int(i): pts[off[k]:off[k + 1]] for k, i in enumerate(ids)}
I6 = oracle.transform_identity()

def check(name, srcs, dsts, T0s, max_iter, dim=2):
    so = np.zeros(len(srcs) + 1, np.uint32); so[1:] = np.cumsum([len(s) for s in srcs])
    do = np.zeros(len(dsts) + 1, np.uint32); do[1:] = np.cumsum([len(d) for d in dsts])
    src = np.concatenate(srcs); dst = np.concatenate(dsts)
    t = time.time()
    g = ctx.estimate_batch(src, so, dst, do, np.array(T0s), max_iter, dim=dim, corr_mode=icp.CORR_ALL)
    dt = time.time() - t
    bad = 0
    for b in range(len(srcs)):
        o = oracle.icp_estimate(srcs[b], dsts[b], T0s[b], max_iter, nn_mode=oracle.NN_KDTREE, corr_mode=oracle.CORR_ALL)
        n = len(srcs[b])
        gc = g["corr"][int(so[b]) * max_iter:int(so[b]) * max_iter + n * max_iter].reshape(max_iter, n)
        ok = (np.array_equal(gc, o["corr"]) and np.array_equal(g["inner"][b], o["inner"])
              and np.array_equal(g["T"][b], o["T"]) and g["corr_hash"][b] == o["corr_hash"])
        if not ok:
            bad += 1
            print("  MISMATCH pair", b, "corr", np.array_equal(gc, o["corr"]), "inner", np.array_equal(g["inner"][b], o["inner"]),
                  "T", np.abs(g["T"][b] - o["T"]).max())
    print(f"{name}: pairs={len(srcs)} bad={bad} gpu_call={dt*1e3:.1f} ms info={ctx.call_info()}")
    return bad

bad = 0
bad += check("scans 001->002..009", [sc[1]] * 8, [sc[k] for k in range(2, 10)], [I6] * 8, 20)
bad += check("scans dup-origin 098->099..102", [sc[98]] * 4, [sc[k] for k in range(99, 103)], [I6] * 4, 5)
src, dst, offs, tt = icp.synth_scan_pairs(64)
bad += check("synthetic 64", [src[i * 1080:(i + 1) * 1080] for i in range(64)], [dst[i * 1080:(i + 1) * 1080] for i in range(64)], [I6] * 64, 20)
print("TOTAL BAD", bad)

# rough throughput, host path
for B in (512, 4096):
    src, dst, offs, tt = icp.synth_scan_pairs(B)
    T0 = np.tile(I6, (B, 1))
    for rep in range(2):
        t = time.time()
        g = ctx.estimate_batch(src, offs, dst, offs, T0, 20)
        dt = time.time() - t
        info = ctx.call_info()
        print(f"B={B} rep={rep} wall={dt*1e3:.1f} ms -> {B/dt:.0f} align/s; kernel_ms(sum)={info['kernel_ms']:.1f} grid={info['grid_ctas']} smem={info['smem_bytes']}")
    ev = (g["inner"] & 0xffff).sum(1) + 20
    print("  evals/pair mean", ev.mean(), "max", ev.max())
