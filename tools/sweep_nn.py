"""GPU sweep of the stand-alone correspondence kernel (nn_search_kernel) on the config-3 map:
cells per point of the shared-target grid x query shape (scan-shaped / map points + jitter).
Usage (on the GPU box):  python tools/sweep_nn.py [cells_per_point ...]   (0 = ring search)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import icp_rust_b200 as icp
    import oracle
    cpps = [int(a) for a in sys.argv[1:]] or [0, 8, 16, 32, 64, 128]
    dev = torch.device("cuda:0")
    ctx = icp.Context(0)
    stream = torch.cuda.Stream(dev)
    m = icp.synth_map(100_000, seed=7)
    md = torch.from_numpy(m).to(dev)
    # scan-shaped queries: config 3 (scans from random poses, perturbed initial transform)
    nscan = 15534
    scans, off, T0 = icp.synth_map_scans(nscan, beams=1080, seed=7)
    sd = torch.from_numpy(scans).to(dev).reshape(nscan, 1080, 2)
    Td = torch.from_numpy(T0).to(dev)
    sx_, sy_ = sd[..., 0], sd[..., 1]
    q_scan = torch.stack([Td[:, None, 0] * sx_ + Td[:, None, 1] * sy_ + Td[:, None, 4],
                          Td[:, None, 2] * sx_ + Td[:, None, 3] * sy_ + Td[:, None, 5]], dim=-1).reshape(-1, 2).contiguous()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    nq = q_scan.shape[0]
    pick = torch.randint(0, len(m), (nq,), device=dev, generator=g)
    q_jit = (md[pick] + torch.randn((nq, 2), dtype=torch.float64, device=dev, generator=g) * 20.0).contiguous()
    out = torch.empty(nq, dtype=torch.int32, device=dev)
    ref = {}
    for cpp in cpps:
        if cpp == 0:
            os.environ["ICP_B200_TARGET_FAST"] = "0"
            os.environ["ICP_B200_TARGET_CELLS_PER_POINT"] = "2"
        else:
            os.environ["ICP_B200_TARGET_FAST"] = "1"
            os.environ["ICP_B200_TARGET_CELLS_PER_POINT"] = str(cpp)
        t0 = time.perf_counter()
        tgt = icp.Target(m, ctx)
        build_ms = (time.perf_counter() - t0) * 1e3
        row = {"cells_per_point": cpp, "mode": "ring" if cpp == 0 else "seeded disc + f32 filter", "build_ms": round(build_ms, 2)}
        for name, q in (("scan", q_scan), ("jitter", q_jit)):
            for _ in range(2):
                tgt.search_device(q.data_ptr(), nq, out.data_ptr(), stream=stream.cuda_stream, synchronize=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                tgt.search_device(q.data_ptr(), nq, out.data_ptr(), stream=stream.cuda_stream, synchronize=False)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            o = out.cpu().numpy()
            if name not in ref:
                ref[name] = o.copy()
                sl = slice(0, 4096)
                ok = bool(np.array_equal(o[sl], oracle.nn_search(m, q[sl].cpu().numpy(), oracle.NN_BRUTE)))
            else:
                ok = bool(np.array_equal(o, ref[name]))
            row[name] = {"ms": round(ms, 3), "Gq_per_s": round(nq / ms / 1e6, 3), "GB_per_s": round(20 * nq / ms / 1e6, 1),
                         "same_indices": ok}
        tgt.close()
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
