#!/usr/bin/env python
"""Headless counterpart of the reference's examples/scan3d.rs: frames of 75 packets from the bundled
3-D recording, nearest neighbours in xyz, SE(2) alignment on xy, warm start, 20 iterations.

    python examples/scan3d.py [--hdf5 scans/3d/scans.hdf5] [--count N]
Without --hdf5 the committed first three frames (tests/golden/scans3d.npz) are used."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from icp_rust_b200 import scans as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hdf5")
    ap.add_argument("--count", type=int, default=None)
    a = ap.parse_args()
    frames = S.load_frames3d_hdf5(a.hdf5) if a.hdf5 else S.load_fixture("scans3d.npz")
    # scan3d.rs:106-121: src = frame 0, and the first dst is frame 0 again
    transforms, path = S.scan_chain([frames[0]] + list(frames), 20, count=a.count)
    print("frame,x,y")
    for k, p in enumerate(path):
        print(f"{k},{p[0]:.9f},{p[1]:.9f}")


if __name__ == "__main__":
    main()
