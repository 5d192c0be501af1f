"""Profiling driver: a few launches of the fused kernel (device-resident inputs) and of the
stand-alone NN kernel, sized for ncu replays.  usage: profile_fused.py [pairs] [launches] [nn_queries]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import icp_rust_b200 as icp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 888
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
NQ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
ctx = icp.Context(0)
src, dst, off, _ = icp.synth_scan_pairs(B, beams=1080)
d = dict(src=torch.from_numpy(src).to(dev), dst=torch.from_numpy(dst).to(dev), off=torch.from_numpy(off.astype(np.int32)).to(dev))
T0 = torch.zeros((B, 6), dtype=torch.float64, device=dev); T0[:, 0] = 1; T0[:, 3] = 1
out = torch.zeros((B, 6), dtype=torch.float64, device=dev)
st = torch.cuda.Stream(device=dev)
for _ in range(L):
    ctx.estimate_batch_device(dict(src=d["src"].data_ptr(), src_off=d["off"].data_ptr(), dst=d["dst"].data_ptr(), dst_off=d["off"].data_ptr(),
                                   init_T=T0.data_ptr(), out_T=out.data_ptr()), B, 20, dim=2, max_src=1080, max_dst=1080,
                              stream=st.cuda_stream, synchronize=True)
    print("fused", B, "pairs:", ctx.call_info())
if NQ:
    m = icp.synth_map(100_000, seed=7)
    tgt = icp.Target(m, ctx)
    g = torch.Generator(device=dev); g.manual_seed(1234)
    md = torch.from_numpy(m).to(dev)
    q = (md[torch.randint(0, 100_000, (NQ,), device=dev, generator=g)] + torch.randn((NQ, 2), dtype=torch.float64, device=dev, generator=g) * 20.0).contiguous()
    o = torch.empty(NQ, dtype=torch.int32, device=dev)
    for _ in range(L):
        tgt.search_device(q.data_ptr(), NQ, o.data_ptr(), stream=st.cuda_stream, synchronize=True)
        print("nn", NQ, ctx.call_info())
