"""Multi-GPU sharding of a batch of scan pairs.

Each `estimate` call of the reference is independent (src/lib.rs:105-130 reads only its own src,
dst and initial_transform), so a batch shards as contiguous slices of the pair index, one slice
per rank / GPU.  There is NO collective on the data path; only the 48-byte result per pair (plus
optional iteration counts and checksums) is gathered afterwards.  One process per GPU,
torch.distributed for the plumbing (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def pair_slice(total_pairs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [begin, end) of rank `rank`: sizes differ by at most one pair and the
    slices tile [0, total_pairs) in rank order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(total_pairs), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def slice_ragged(points: np.ndarray, off: np.ndarray, begin: int, end: int):
    """Rows of a concatenated ragged point array belonging to pairs [begin, end), with offsets
    rebased to zero (what one rank passes to icp_b200_estimate_batch)."""
    off = np.asarray(off, dtype=np.uint32)
    p0, p1 = int(off[begin]), int(off[end])
    return points[p0:p1], (off[begin:end + 1] - off[begin]).astype(np.uint32)


def gather_rows(local: np.ndarray, total_pairs: int, group=None, device=None):
    """All-gather per-pair result rows (e.g. out_T (b,6), inner (b,K), corr_hash (b,)) from every
    rank's slice into the full (total_pairs, ...) array, in pair order, on every rank.
    Slices may differ by one row, so rows are padded to the largest slice for the collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b0, b1 = pair_slice(total_pairs, rank, world)
    local = np.ascontiguousarray(local)
    assert local.shape[0] == b1 - b0, "local rows do not match this rank's slice"
    cap = -(-total_pairs // world)
    # reinterpret as bytes so that u64 checksums survive backends without unsigned support
    row_bytes = local.dtype.itemsize * int(np.prod(local.shape[1:], dtype=np.int64))
    buf = np.zeros((cap, row_bytes), dtype=np.uint8)
    buf[:b1 - b0] = local.reshape(b1 - b0, -1).view(np.uint8).reshape(b1 - b0, row_bytes)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    parts = []
    for r, o in enumerate(outs):
        s0, s1 = pair_slice(total_pairs, r, world)
        parts.append(o.cpu().numpy()[:s1 - s0])
    full = np.concatenate(parts, axis=0)
    return full.view(local.dtype).reshape((total_pairs,) + local.shape[1:])
