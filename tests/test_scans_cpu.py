"""Host-side pieces of the scan-chain driver and the multi-hypothesis front end (no GPU):
loaders as the reference's examples parse their recordings, Transform::inverse."""
import os

import numpy as np

import oracle
from icp_rust_b200 import Transform
from icp_rust_b200 import scans as S


def test_scan2d_text_loader(tmp_path):  # examples/scan2d.rs:18-34
    p = tmp_path / "001.txt"
    p.write_text("1.5 -2.25\n0 0\n-3e2 4.0e-1\n")
    (tmp_path / "002.txt").write_text("7 8\n")
    (tmp_path / "notes.md").write_text("ignored")
    a = S.load_scan2d_txt(str(p))
    assert a.dtype == np.float64 and np.array_equal(a, [[1.5, -2.25], [0., 0.], [-300., 0.4]])
    both = S.load_scans2d_dir(str(tmp_path))
    assert len(both) == 2 and np.array_equal(both[1], [[7., 8.]])


def test_fixture_matches_recording_shape():
    scans = S.load_fixture("scans2d.npz")
    assert len(scans) == 210 and all(s.shape[1] == 2 and len(s) > 100 for s in scans)
    frames = S.load_fixture("scans3d.npz")
    assert len(frames) == 3 and all(f.shape[1] == 3 for f in frames)
    # the committed frames already passed the example's norm filter (scan3d.rs:63-69)
    assert all(len(S.remove_invalid_values(f)) == len(f) for f in frames)
    assert len(S.remove_invalid_values([[0.1, 0.1, 0.1], [0.2, 0., 0.], [1., 0., 0.]])) == 1


def test_inverse_follows_the_reference():
    T = oracle.transform_new([120.0, -35.0, 0.3])
    Ti = Transform.from_array(T).inverse()                       # transform.rs:26-32
    assert np.allclose(Ti.to_array(), oracle.transform_inverse(T), rtol=0, atol=1e-12)


def test_examples_exist_and_parse():
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    for f in ("examples/scan2d.py", "examples/scan3d.py"):
        compile(open(os.path.join(root, f)).read(), f, "exec")
