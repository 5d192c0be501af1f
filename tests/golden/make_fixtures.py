"""Regenerates tests/golden/*.npz from the reference's bundled scans.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
The outputs are committed; tests and bench.py read only the committed files.

scans2d.npz : every file of /root/reference/scans/2d (000..209), parsed exactly as
              examples/scan2d.rs:18-34 does (one "x y" per line, split on a single space,
              f64).  Stored as one concatenated (P,2) float64 array plus offsets.
scans3d.npz : the first three frames of /root/reference/scans/3d/scans.hdf5 as examples/scan3d.rs
              builds them: 75 packets of (24,16,3) f64 per frame (:9,:104), reshaped to (384,3)
              (:21-23), concatenated (:49-61), points with norm <= 0.2 removed (:63-69).
              Read with the dependency-free reader tests/golden/hdf5_min.py (no h5py here).
scans3d_20.npz : all 20 frames of the recording (examples/scan3d.rs:104-131 walks them), same
              framing and filter, cut to the FIRST 16,384 points of each frame: the clouds of
              BASELINE.json config 4 (1024 pairs (frame i, frame i+1 mod 20) x 16,384 points).
"""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_scan2d(path):
    pts = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            xy = line.split(" ")
            pts.append((float(xy[0]), float(xy[1])))
    return np.array(pts, dtype=np.float64).reshape(-1, 2)


def frames3d(n_frames=3, step=75, cut=None, name="scans3d.npz"):
    sys.path.insert(0, HERE)
    from hdf5_min import MiniH5
    h = MiniH5(os.path.join(REF, "scans", "3d", "scans.hdf5"))
    ds = h.datasets()
    names = list(ds)                       # zero-padded names: lexicographic = numeric order
    assert len(names) == 1500
    frames = []
    for k in range(n_frames):
        pts = np.concatenate([h.read(ds[names[i]]).reshape(24 * 16, 3) for i in range(step * k, step * (k + 1))])
        nrm = np.sqrt(pts[:, 0] * pts[:, 0] + pts[:, 1] * pts[:, 1] + pts[:, 2] * pts[:, 2])  # norm.rs:8-21
        frames.append(pts[nrm > 0.2][:cut] if cut else pts[nrm > 0.2])
    off = np.zeros(n_frames + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(f) for f in frames])
    out = os.path.join(HERE, name)
    np.savez_compressed(out, points=np.concatenate(frames), offsets=off)
    print(out, [f.shape for f in frames], os.path.getsize(out))


def main():
    frames3d()
    frames3d(n_frames=20, cut=16384, name="scans3d_20.npz")
    d = os.path.join(REF, "scans", "2d")
    names = sorted(f for f in os.listdir(d) if f.endswith(".txt"))
    scans = [load_scan2d(os.path.join(d, n)) for n in names]
    off = np.zeros(len(scans) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in scans])
    pts = np.concatenate(scans, axis=0)
    ids = np.array([int(n[:3]) for n in names], dtype=np.int32)
    out = os.path.join(HERE, "scans2d.npz")
    np.savez_compressed(out, points=pts, offsets=off, ids=ids)
    print(out, pts.shape, os.path.getsize(out))


if __name__ == "__main__":
    sys.exit(main())
