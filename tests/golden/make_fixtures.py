"""Regenerates tests/golden/*.npz from the reference's bundled scans.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
The outputs are committed; tests and bench.py read only the committed files.

scans2d.npz : every file of /root/reference/scans/2d (000..209), parsed exactly as
              examples/scan2d.rs:18-34 does (one "x y" per line, split on a single space,
              f64).  Stored as one concatenated (P,2) float64 array plus offsets.
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


def main():
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
