#!/usr/bin/env python
"""Headless counterpart of the reference's examples/scan2d.rs: the bundled 2-D LiDAR recording
aligned scan by scan (src = first scan, warm start, 20 iterations), sensor path printed as CSV.

    python examples/scan2d.py [--scans-dir scans/2d] [--count N]
Without --scans-dir the committed copy of the recording (tests/golden/scans2d.npz) is used."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from icp_rust_b200 import scans as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans-dir")
    ap.add_argument("--count", type=int, default=None, help="number of alignments (default: the whole recording)")
    a = ap.parse_args()
    scans = S.load_scans2d_dir(a.scans_dir) if a.scans_dir else S.load_fixture("scans2d.npz")[1:]   # 001.txt is the first the example reads
    transforms, path = S.scan_chain(scans, 20, count=a.count)
    print("index,x,y,r00,r01,r10,r11")
    for k, (T, p) in enumerate(zip(transforms, path)):
        r = T.inverse().rot
        print(f"{k + 2},{p[0]:.9f},{p[1]:.9f},{r[0, 0]:.12f},{r[0, 1]:.12f},{r[1, 0]:.12f},{r[1, 1]:.12f}")


if __name__ == "__main__":
    main()
