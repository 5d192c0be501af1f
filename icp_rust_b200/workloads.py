"""Workload builders for BASELINE.json configs 3-5 (shared by tests/, bench.py and tools/).

Inputs only: nothing here computes an alignment.  Config 2's generator is `synth_scan_pairs`
(csrc/synth.c), config 3's `synth_map` / `synth_map_scans`; config 4 is built from the committed
copy of the reference's 3-D recording (tests/golden/scans3d_20.npz, made by
tests/golden/make_fixtures.py from /root/reference/scans/3d/scans.hdf5 exactly as
examples/scan3d.rs:9,21-69,104 frames it)."""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG4_POINTS = 16384
CONFIG4_FRAMES = 20


def load_frames3d_20():
    """The 20 frames of the bundled 3-D recording, first 16,384 filtered points each."""
    z = np.load(os.path.join(_HERE, "..", "tests", "golden", "scans3d_20.npz"))
    pts, off = z["points"], z["offsets"]
    return [pts[off[k]:off[k + 1]] for k in range(len(off) - 1)]


def config4_init_transforms(count, seed=4, first_pair=0):
    """Seeded SE(2) initial perturbations U(+-0.02 m, +-0.02 m, +-0.01 rad);
    pair p always gets the same transform whatever `count` / `first_pair` slice is asked for."""
    T0 = np.empty((count, 6))
    for k in range(count):
        rng = np.random.default_rng([seed, first_pair + k])
        tx, ty = rng.uniform(-0.02, 0.02, 2)
        e = rng.uniform(-0.01, 0.01)
        T0[k] = [np.cos(e), -np.sin(e), np.sin(e), np.cos(e), tx, ty]
    return T0


def config4_pairs(count, seed=4, first_pair=0, frames=None, out_src=None, out_dst=None):
    """BASELINE.json config 4: pair p = (frame p mod 20, frame (p+1) mod 20) of the bundled 3-D
    recording (examples/scan3d.rs:104-131 walks the same frames), 16,384 points per cloud, each
    pair with its own seeded initial transform.  Returns (src (count*16384,3), dst, off, T0)."""
    frames = frames if frames is not None else load_frames3d_20()
    n = CONFIG4_POINTS
    src = out_src if out_src is not None else np.empty((count * n, 3))
    dst = out_dst if out_dst is not None else np.empty((count * n, 3))
    for k in range(count):
        p = first_pair + k
        src[k * n:(k + 1) * n] = frames[p % len(frames)][:n]
        dst[k * n:(k + 1) * n] = frames[(p + 1) % len(frames)][:n]
    off = (np.arange(count + 1, dtype=np.uint64) * n).astype(np.uint32)
    return src, dst, off, config4_init_transforms(count, seed, first_pair)
