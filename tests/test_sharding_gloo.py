"""Multi-GPU sharding logic on CPU: world_size-2 gloo processes, one pair slice per rank, only
results gathered (icp_rust_b200/shard.py).  The per-rank compute here is the ORACLE (this is a test
of the plumbing, the GPU kernels are covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

from icp_rust_b200.shard import pair_slice, slice_ragged

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pair_slice_tiles_the_batch():
    for total in (0, 1, 7, 8, 4096, 1_000_000):
        for world in (1, 2, 3, 8):
            edges = [pair_slice(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pair_slice(10, 2, 2)


def test_slice_ragged_rebases_offsets():
    off = np.array([0, 3, 3, 10, 12], dtype=np.uint32)
    pts = np.arange(24, dtype=np.float64).reshape(12, 2)
    p, o = slice_ragged(pts, off, 1, 3)
    assert np.array_equal(o, [0, 0, 7]) and np.array_equal(p, pts[3:10])


def _worker(rank, world, port, total, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import icp_rust_b200 as icp
    import oracle
    from icp_rust_b200.shard import gather_rows, pair_slice

    dist.init_process_group("gloo", rank=rank, world_size=world)
    b0, b1 = pair_slice(total, rank, world)
    beams = 60
    # each rank generates ONLY its slice of the global pair stream
    src, dst, off, _ = icp.synth_scan_pairs(b1 - b0, beams=beams, first_pair=b0, nthreads=1)
    T0 = np.tile(np.array([1., 0., 0., 1., 0., 0.]), (b1 - b0, 1))
    o = oracle.icp_estimate_batch(2, src, off, dst, off, T0, 3, nthreads=1)
    allT = gather_rows(o["T"], total)
    allH = gather_rows(o["corr_hash"], total)
    allI = gather_rows(o["inner"], total)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        out_q.put((allT, allH, allI))


def test_two_rank_gloo_gather_equals_single_process():
    import torch.multiprocessing as mp

    import icp_rust_b200 as icp
    import oracle
    total, world = 7, 2      # uneven split: 4 + 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    allT, allH, allI = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    src, dst, off, _ = icp.synth_scan_pairs(total, beams=60, first_pair=0, nthreads=1)
    T0 = np.tile(np.array([1., 0., 0., 1., 0., 0.]), (total, 1))
    ref = oracle.icp_estimate_batch(2, src, off, dst, off, T0, 3, nthreads=1)
    assert np.array_equal(allT, ref["T"]) and np.array_equal(allH, ref["corr_hash"]) and np.array_equal(allI, ref["inner"])
