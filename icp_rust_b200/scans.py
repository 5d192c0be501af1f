"""Scan loaders and the scan-chain driver of the reference's examples, headless.

examples/scan2d.rs:18-34,62-105 and examples/scan3d.rs:9,21-69,104-148 load a recording, keep the
FIRST scan as `src`, and for every following scan build `IcpXd::new(dst)` and call
`estimate(src, &transform, 20)` warm-started with the previous result; the sensor path is the
translation of `transform.inverse()` (src/transform.rs:26-32).  `scan_chain` does exactly that
through this package's `Icp2d` / `Icp3d` (i.e. through the C ABI and the CUDA kernels); the GUI
drawing of the examples is out of scope.
"""
from __future__ import annotations

import os

import numpy as np



def load_scan2d_txt(path):
    """One 2-D scan as examples/scan2d.rs:18-34 parses it: a point per line, "x y" split on one space."""
    pts = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            xy = line.split(" ")
            pts.append((float(xy[0]), float(xy[1])))
    return np.array(pts, dtype=np.float64).reshape(-1, 2)


def load_scans2d_dir(directory):
    """scans/2d/001.txt, 002.txt, ... in numeric order (examples/scan2d.rs:64 formats {:03}.txt)."""
    names = sorted(f for f in os.listdir(directory) if f.endswith(".txt"))
    return [load_scan2d_txt(os.path.join(directory, n)) for n in names]


def remove_invalid_values(points, min_norm=0.2):
    """examples/scan3d.rs:63-69: keep points with icp::norm(p) > 0.2 (norm.rs:8-21)."""
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    nrm = np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2])
    return p[nrm > min_norm]


def load_frames3d_hdf5(path, step=75, max_frames=None):
    """Frames of scans/3d/scans.hdf5 as examples/scan3d.rs builds them: `step` packets of (24,16,3)
    f64 per frame (:9,:104), reshaped to (384,3) (:21-23) and concatenated (:49-61), invalid returns
    removed.  Uses the package's dependency-free reader (hdf5_min.py)."""
    from .hdf5_min import MiniH5
    h = MiniH5(path)
    ds = h.datasets()
    names = list(ds)   # zero-padded names: lexicographic = numeric order
    frames = []
    k = 0
    while (k + 1) * step <= len(names) and (max_frames is None or k < max_frames):
        pts = np.concatenate([h.read(ds[names[i]]).reshape(-1, 3) for i in range(step * k, step * (k + 1))])
        frames.append(remove_invalid_values(pts))
        k += 1
    return frames


def load_fixture(name):
    """tests/golden/scans2d.npz / scans3d.npz (committed copies of the bundled recordings) -> list of scans."""
    here = os.path.dirname(os.path.abspath(__file__))
    z = np.load(os.path.join(here, "..", "tests", "golden", name))
    pts, off = z["points"], z["offsets"]
    return [pts[off[k]:off[k + 1]] for k in range(len(off) - 1)]


def scan_chain(scans, max_iter=20, ctx=None, first=0, count=None, ordered=True):
    """The loop of examples/scan2d.rs:62-105 / scan3d.rs:104-148 without the window.
    Returns (transforms, path): `transforms[k]` aligns scans[first] to scans[first+1+k], warm-started
    from transforms[k-1]; `path[k]` is the translation of its inverse."""
    from . import Icp2d, Icp3d, Transform
    scans = [np.ascontiguousarray(s, dtype=np.float64) for s in scans]
    dim = scans[first].shape[1]
    cls = Icp2d if dim == 2 else Icp3d
    src = scans[first]
    T = Transform.identity()
    transforms, path = [], []
    last = len(scans) if count is None else min(len(scans), first + 1 + count)
    for k in range(first + 1, last):
        T = cls.new(scans[k], ctx).estimate(src, T, max_iter, ordered=ordered)   # ordered: the reference's own bits
        transforms.append(T)
        path.append(T.inverse().t.copy())
    return transforms, np.array(path).reshape(-1, 2)


def estimate_best_of(src, dst, initial_transforms, max_iter=20, ctx=None):
    """Multi-hypothesis front end: align `src` to `dst` from every initial transform
    in ONE batched call — one device target (Icp2d::new once), ONE copy of the scan
    (ICP_B200_FLAG_SHARED_SOURCE), H initial transforms — and score each result with the cost the
    loop itself minimises: `huber_error` (lib.rs:45-50) of the last inner solve over the last
    iteration's correspondences, which the kernel returns in out_cost.  The best is an argmin over
    that array.  2-D or 3-D.
    Returns dict(best=index, T=(H,6) results, cost=(H,) scores, transform=Transform of the best)."""
    from . import Target, Transform, default_context
    ctx = ctx or default_context()
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.ascontiguousarray(dst, dtype=np.float64)
    T0 = np.ascontiguousarray([t.to_array() if isinstance(t, Transform) else np.asarray(t, dtype=np.float64).reshape(6)
                               for t in initial_transforms]).reshape(-1, 6)
    tgt = Target(dst, ctx)
    try:
        res = tgt.estimate_batch(src, np.array([0, src.shape[0]], dtype=np.uint32), T0, max_iter, want_inner=False,
                                 want_hash=False, want_cost=True, shared_source=True)
    finally:
        tgt.close()
    cost = res["cost"]
    best = int(np.argmin(cost))
    return {"best": best, "T": res["T"], "cost": cost, "transform": Transform.from_array(res["T"][best]),
            "status": res["status"]}
