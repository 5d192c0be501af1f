"""Host timeline of the pipelined submit / wait loop on bench.py's buffers of rank R (R = argv[1]): when each
submit and wait starts and ends, per step.  Diagnosis of what a guard re-run costs a pipelining caller."""
import sys
import time

import numpy as np
import torch

import icp_rust_b200 as icp

R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
B, BEAMS = 4096, 1080
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cs = [icp.Context(0), icp.Context(0)]
host = []
for j in range(2):
    hs = torch.empty((B * BEAMS, 2), dtype=torch.float64, pin_memory=True)
    hd = torch.empty((B * BEAMS, 2), dtype=torch.float64, pin_memory=True)
    _, _, off, _ = icp.synth_scan_pairs(B, beams=BEAMS, first_pair=(R * 2 + j) * B, out_src=hs.numpy(), out_dst=hd.numpy())
    hT0 = torch.zeros((B, 6), dtype=torch.float64); hT0[:, 0] = 1.0; hT0[:, 3] = 1.0
    host.append(dict(src=hs, dst=hd, off=torch.from_numpy(off.astype(np.int32)).pin_memory(), T0=hT0.pin_memory(),
                     out=torch.zeros((B, 6), dtype=torch.float64).pin_memory(), status=torch.zeros(B, dtype=torch.int32).pin_memory()))


def submit(j):
    h = host[j % 2]
    cs[j % 2].submit_host_ptrs(dict(src=h["src"].data_ptr(), src_off=h["off"].data_ptr(), dst=h["dst"].data_ptr(),
                                    dst_off=h["off"].data_ptr(), init_T=h["T0"].data_ptr(), out_T=h["out"].data_ptr(),
                                    out_status=h["status"].data_ptr()), B, 20, dim=2)


for w in range(2):   # warm-up
    submit(0); submit(1); cs[0].wait(); cs[1].wait()
torch.cuda.synchronize()
t0 = time.perf_counter()
ev = []
now = lambda: (time.perf_counter() - t0) * 1e3
a = now(); submit(0); ev.append(("submit", 0, a, now()))
for j in range(1, K):
    a = now(); submit(j); b = now(); ev.append(("submit", j, a, b))
    a = now(); cs[(j - 1) % 2].wait(); b = now(); ev.append(("wait", j - 1, a, b, cs[(j - 1) % 2].call_info()["redo_pairs"], round(cs[(j - 1) % 2].call_info()["kernel_ms"], 2)))
a = now(); cs[(K - 1) % 2].wait(); b = now(); ev.append(("wait", K - 1, a, b, cs[(K - 1) % 2].call_info()["redo_pairs"]))
for e in ev:
    print(" ".join(str(round(x, 2)) if isinstance(x, float) else str(x) for x in e))
print("ms per step", round(now() / K, 2))
