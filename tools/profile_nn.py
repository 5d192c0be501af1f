"""Launches nn_search_kernel on the config-3 map for ncu: 2 warm-up launches, then one launch on
scan-shaped queries and one on jittered map points (capture with -k regex:nn_search -s 2 -c 2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import icp_rust_b200 as icp
    nscan = int(sys.argv[1]) if len(sys.argv) > 1 else 15534
    dev = torch.device("cuda:0")
    ctx = icp.Context(0)
    m = icp.synth_map(100_000, seed=7)
    md = torch.from_numpy(m).to(dev)
    scans, off, T0 = icp.synth_map_scans(nscan, beams=1080, seed=7)
    sd = torch.from_numpy(scans).to(dev).reshape(nscan, 1080, 2)
    Td = torch.from_numpy(T0).to(dev)
    sx_, sy_ = sd[..., 0], sd[..., 1]
    q_scan = torch.stack([Td[:, None, 0] * sx_ + Td[:, None, 1] * sy_ + Td[:, None, 4],
                          Td[:, None, 2] * sx_ + Td[:, None, 3] * sy_ + Td[:, None, 5]], dim=-1).reshape(-1, 2).contiguous()
    nq = q_scan.shape[0]
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    pick = torch.randint(0, len(m), (nq,), device=dev, generator=g)
    q_jit = (md[pick] + torch.randn((nq, 2), dtype=torch.float64, device=dev, generator=g) * 20.0).contiguous()
    out = torch.empty(nq, dtype=torch.int32, device=dev)
    tgt = icp.Target(m, ctx)
    torch.cuda.synchronize()
    for q in (q_scan, q_jit, q_scan, q_jit):
        tgt.search_device(q.data_ptr(), nq, out.data_ptr(), synchronize=True)
    tgt.close()


if __name__ == "__main__":
    main()
