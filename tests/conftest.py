import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def scans2d():
    """Bundled 2-D LiDAR scans of the reference (tests/golden/scans2d.npz, see make_fixtures.py)."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "scans2d.npz"))
    pts, off, ids = z["points"], z["offsets"], z["ids"]
    return {int(i): pts[off[k]:off[k + 1]].copy() for k, i in enumerate(ids)}


@pytest.fixture(scope="session")
def scans3d():
    """First three bundled 3-D frames (tests/golden/scans3d.npz, see make_fixtures.py)."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "scans3d.npz"))
    pts, off = z["points"], z["offsets"]
    return [pts[off[k]:off[k + 1]].copy() for k in range(len(off) - 1)]


@pytest.fixture(scope="session")
def frames3d_20():
    """All 20 bundled 3-D frames cut to 16,384 points (tests/golden/scans3d_20.npz): config 4."""
    from icp_rust_b200.workloads import load_frames3d_20
    return load_frames3d_20()
