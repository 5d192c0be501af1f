"""Timing of the config-4 shape (3-D, 16,384-point clouds cut from the bundled frames) on the
streaming kernel.  usage: time_config4.py [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import icp_rust_b200 as icp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scans3d.npz"))
pts, off = z["points"], z["offsets"]
fr = [pts[off[k]:off[k + 1]][:16384] for k in range(3)]
rng = np.random.default_rng(4)
srcs, dsts, T0 = [], [], []
for b in range(B):
    i, j = [(0, 1), (1, 2), (0, 2), (1, 0), (2, 1), (2, 0)][b % 6]
    th = rng.uniform(-3, 3); c, s = np.cos(th), np.sin(th); t = rng.uniform(-1, 1, 2)
    R = np.array([[c, -s], [s, c]])
    mv = lambda p: np.column_stack([p[:, :2] @ R.T + t, p[:, 2]])
    srcs.append(mv(fr[i])); dsts.append(mv(fr[j]))
    e = rng.uniform(-0.01, 0.01); T0.append([np.cos(e), -np.sin(e), np.sin(e), np.cos(e), rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02)])
src = np.concatenate(srcs); dst = np.concatenate(dsts)
o = (np.arange(B + 1, dtype=np.uint64) * 16384).astype(np.uint32)
ctx = icp.Context(0)
for rep in range(3):
    t = time.time()
    g = ctx.estimate_batch(src, o, dst, o, np.array(T0), 20, dim=3)
    dt = time.time() - t
    info = ctx.call_info()
    print(f"config4 shape: {B} pairs x 16384 pts: wall {dt*1e3:.1f} ms, kernel {info['kernel_ms']:.1f} ms -> {B/(info['kernel_ms']*1e-3):.0f} pairs/s (kernel), {B*16384*20/(info['kernel_ms']*1e-3)/1e9:.2f} G NN queries/s incl. solve; grid {info['grid_ctas']}x{info['block_threads']}")
print("accepted per outer iteration:", (g["inner"] & 0xffff).mean(), "status ok:", (g["status"] == 0).all())

# CPU oracle on the first pairs of the same batch (exact kd-tree, one pair per host thread): parity at
# this size and a baseline for the shape
import oracle
nchk = min(B, os.cpu_count() or 1)
oo = (np.arange(nchk + 1, dtype=np.uint64) * 16384).astype(np.uint32)
t = time.time()
o = oracle.icp_estimate_batch(3, src[:nchk * 16384], oo, dst[:nchk * 16384], oo, np.array(T0)[:nchk], 20,
                              nn_mode=oracle.NN_KDTREE, nthreads=nchk)
dt = time.time() - t
print(f"oracle: {nchk} pairs on {nchk} host threads in {dt:.2f} s -> {nchk / dt:.1f} pairs/s; "
      f"max |T_gpu - T_oracle| = {np.abs(o['T'] - g['T'][:nchk]).max():.3e}; "
      f"inner counts equal: {bool(np.array_equal(o['inner'], g['inner'][:nchk]))}; "
      f"correspondence hashes equal: {bool(np.array_equal(o['corr_hash'], g['corr_hash'][:nchk]))}")
