"""Small invocations of every kernel path added late in round 2, sized for compute-sanitizer
(racecheck / memcheck slow a launch down ~100x): the lean and ordered kernels on bundled 2-D
scans that reach the outer loop's fixed point, the streaming kernel with the binary32 record
filter on subsampled 3-D frames, back-to-back asynchronous device calls with
ICP_B200_FLAG_INDEPENDENT (and every pair forced through the re-run), the group search kernel.
Every result is compared with the oracle.  usage: sanitizer_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import icp_rust_b200 as icp
import oracle
from icp_rust_b200.workloads import load_frames3d_20, config4_init_transforms

I6 = np.array([1., 0., 0., 1., 0., 0.])


def cat(arrs, dim):
    off = np.zeros(len(arrs) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(a) for a in arrs])
    return np.concatenate(arrs).reshape(-1, dim).astype(np.float64), off


def main():
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scans2d.npz"))
    pts, off = z["points"], z["offsets"]
    scans = [pts[off[k]:off[k + 1]] for k in range(8)]
    ctx = icp.Context(0)
    # 2-D: lean + ordered, fixed point reached (bundled scans stop moving early)
    s, so = cat(scans[0:6], 2)
    d, do = cat(scans[1:7], 2)
    T0 = np.tile(I6, (6, 1))
    o = oracle.icp_estimate_batch(2, s, so, d, do, T0, 20, nn_mode=oracle.NN_KDTREE)
    for ordered in (False, True):
        g = ctx.estimate_batch(s, so, d, do, T0, 20, corr_mode=icp.CORR_ALL, want_cost=True, ordered=ordered)
        assert np.array_equal(g["inner"], o["inner"]) and np.array_equal(g["corr_hash"], o["corr_hash"])
        assert np.abs(g["T"] - o["T"]).max() <= (0 if ordered else 1e-9)
    assert ((o["inner"] & 0xffff) == 0)[:, :-1].any(axis=1).any(), "no pair reached the fixed point"
    print("2-D lean / ordered with the fixed-point exit: ok")
    # 3-D streaming kernel (scans beyond the fused instances' capacity), binary32 record filter
    fr = load_frames3d_20()
    s3, o3 = cat([fr[1][:3600], fr[2][:3600]], 3)
    d3, p3 = cat([fr[0][:3600], fr[1][:3600]], 3)
    T3 = config4_init_transforms(2)
    o = oracle.icp_estimate_batch(3, s3, o3, d3, p3, T3, 8, nn_mode=oracle.NN_KDTREE)
    g = ctx.estimate_batch(s3, o3, d3, p3, T3, 8, dim=3)
    assert ctx.call_info()["block_threads"] == 512
    assert np.array_equal(g["inner"], o["inner"]) and np.array_equal(g["corr_hash"], o["corr_hash"])
    assert np.abs(g["T"] - o["T"]).max() <= 1e-9
    print("3-D streaming kernel with the binary32 filter: ok")
    ctx.close()
    # back-to-back asynchronous device calls, every pair through the re-run launch
    os.environ["ICP_B200_GUARD_SCALE"] = "1e9"
    ctx = icp.Context(0)
    os.environ.pop("ICP_B200_GUARD_SCALE")
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    src, dst, offp, _ = icp.synth_scan_pairs(6, beams=360, first_pair=900)
    want = ctx.estimate_batch(src, offp, dst, offp, np.tile(I6, (6, 1)), 6)
    bufs = []
    for k in range(3):
        bufs.append(dict(src=torch.from_numpy(src).to(dev), dst=torch.from_numpy(dst).to(dev),
                         off=torch.from_numpy(offp.astype(np.int32)).to(dev), T0=torch.from_numpy(np.tile(I6, (6, 1))).to(dev),
                         out=torch.zeros((6, 6), dtype=torch.float64, device=dev)))
    torch.cuda.synchronize()
    for rep in range(2):
        for b in bufs:
            ctx.estimate_batch_device(dict(src=b["src"].data_ptr(), src_off=b["off"].data_ptr(), dst=b["dst"].data_ptr(),
                                           dst_off=b["off"].data_ptr(), init_T=b["T0"].data_ptr(), out_T=b["out"].data_ptr()),
                                      6, 6, dim=2, max_src=360, max_dst=360, stream=stream.cuda_stream, synchronize=False,
                                      flags=icp.FLAG_INDEPENDENT)
    torch.cuda.synchronize()
    for b in bufs:
        assert np.array_equal(b["out"].cpu().numpy(), want["T"])
    print("overlapping device calls with forced re-runs: ok")
    ctx.close()
    # group search kernel
    for g_ in ("4", "8"):
        os.environ["ICP_B200_NN_GROUP"] = g_
        ctx = icp.Context(0)
        m = icp.synth_map(4000, seed=7)
        tgt = icp.Target(m, ctx)
        q = m[::3] + np.random.default_rng(1).normal(0, 30, m[::3].shape)
        assert np.array_equal(tgt.search(q), oracle.nn_search(m, q, oracle.NN_BRUTE))
        tgt.close()
        ctx.close()
    os.environ.pop("ICP_B200_NN_GROUP")
    print("group search kernels: ok")


if __name__ == "__main__":
    main()
