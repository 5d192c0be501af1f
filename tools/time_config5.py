"""BASELINE config 5: 1M loop-closure candidate scan pairs (config-2 shape) sharded across the GPUs
of one box — contiguous slices of the pair index, one process per GPU, inputs device-resident when
the clock runs, CUDA-event time per rank, max over ranks; only results are gathered.

  python tools/time_config5.py [--pairs 1000000] [--chunk 16384]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/time_config5.py

Each rank streams its slice through the device in chunks (generate on the host -> H2D -> timed
kernel); the synthetic stream is counter-based, so every world size aligns the same 1M pairs.
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--check", type=int, default=32, help="pairs of rank 0's first chunk compared with the oracle")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import icp_rust_b200 as icp
    from icp_rust_b200.shard import pair_slice
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = icp.Context(local)
    stream = torch.cuda.Stream(dev)
    b0, b1 = pair_slice(a.pairs, rank, world)
    beams, K = 1080, 20
    nthreads = max(1, (os.cpu_count() or 1) // world)
    kernel_ms, gen_s, done = 0.0, 0.0, 0
    checksum = np.zeros(6)
    first = None
    cap = min(a.chunk, b1 - b0)
    d_src = torch.empty((cap * beams, 2), dtype=torch.float64, device=dev)
    d_dst = torch.empty((cap * beams, 2), dtype=torch.float64, device=dev)
    d_off = torch.from_numpy((np.arange(cap + 1, dtype=np.int64) * beams).astype(np.int32)).to(dev)
    d_T0 = torch.zeros((cap, 6), dtype=torch.float64, device=dev)
    d_T0[:, 0] = 1.0
    d_T0[:, 3] = 1.0
    d_out = torch.empty((cap, 6), dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
    wall0 = time.perf_counter()
    for c0 in range(b0, b1, a.chunk):
        n = min(a.chunk, b1 - c0)
        t = time.perf_counter()
        src, dst, off, _ = icp.synth_scan_pairs(n, beams=beams, first_pair=c0, nthreads=nthreads)
        gen_s += time.perf_counter() - t
        d_src[:n * beams].copy_(torch.from_numpy(src))
        d_dst[:n * beams].copy_(torch.from_numpy(dst))
        torch.cuda.synchronize()
        ptrs = dict(src=d_src.data_ptr(), src_off=d_off.data_ptr(), dst=d_dst.data_ptr(), dst_off=d_off.data_ptr(),
                    init_T=d_T0.data_ptr(), out_T=d_out.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.estimate_batch_device(ptrs, n, K, dim=2, max_src=beams, max_dst=beams, stream=stream.cuda_stream, synchronize=False)
        e1.record(stream)
        torch.cuda.synchronize()
        kernel_ms += e0.elapsed_time(e1)
        out = d_out[:n].cpu().numpy()
        checksum += out.sum(axis=0)
        if first is None:
            first = (src, dst, off, out)
        done += n
    wall = time.perf_counter() - wall0
    t_max = torch.tensor([kernel_ms, wall], dtype=torch.float64, device=dev)
    cs = torch.from_numpy(checksum).to(dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    if rank == 0:
        line = {"workload": f"config5: {a.pairs} synthetic 2D scan pairs x {beams} points, {K} outer iterations, "
                            f"contiguous slices over {world} GPU(s), chunks of {a.chunk} pairs",
                "n_gpus": world, "pairs": a.pairs, "device_seconds_max_over_ranks": float(t_max[0]) * 1e-3,
                "pairs_per_s": a.pairs / (float(t_max[0]) * 1e-3),
                "wall_seconds_incl_host_generation_and_copies": float(t_max[1]),
                "result_checksum": [float(v) for v in cs.cpu().numpy()]}
        if a.check:
            import oracle
            src, dst, off, out = first
            n = min(a.check, len(off) - 1)
            worst = 0.0
            for b in range(n):
                o = oracle.icp_estimate(src[off[b]:off[b + 1]], dst[off[b]:off[b + 1]], None, K, nn_mode=oracle.NN_KDTREE)
                worst = max(worst, float(np.abs(o["T"] - out[b]).max()))
            line["oracle_check"] = {"pairs": n, "max_abs_transform_difference": worst}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
