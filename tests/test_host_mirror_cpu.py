"""Host-side mirror of the reference's value types (icp_rust_b200.Transform) against the oracle and
the reference's own transform.rs / se2.rs known answers.  No GPU needed."""
import math

import numpy as np
import pytest

import icp_rust_b200 as icp
import oracle
from tests import refdata as R


def test_transform_known_answers():  # transform.rs:62-96
    T = icp.Transform.from_rt(oracle.so2_new_rotation2(math.pi / 2), [3., 6.])
    assert np.linalg.norm(T.transform([4., 2.]) - np.array([1., 10.])) < 1e-8
    assert np.linalg.norm(T.inverse().transform([1., 10.]) - np.array([4., 2.])) < 1e-8
    T1 = icp.Transform.from_rt(oracle.so2_new_rotation2(math.pi / 4), [2., 1.]).inverse()
    T2 = icp.Transform.from_rt(oracle.so2_new_rotation2(math.pi / 2), [5., 3.]).inverse()
    x = [-5., 6.]
    assert np.linalg.norm(T1.transform(T2.transform(x)) - (T1 * T2).transform(x)) < 1e-8


@pytest.mark.parametrize("param,expected", R.SE2_EXP_CASES)
def test_transform_new_matches_se2_exp(param, expected):  # se2.rs:85-142
    T = icp.Transform.new(param)
    e = np.array(expected)
    assert np.abs(T.rot - e[:2, :2]).max() < 1e-6 and np.abs(T.t - e[:2, 2]).max() < 1e-6
    assert np.abs(T.to_array() - oracle.transform_new(param)).max() < 1e-15


def test_round_trip_and_identity():
    T = icp.Transform.new([1., -2., 0.3])
    assert np.array_equal(icp.Transform.from_array(T.to_array()).to_array(), T.to_array())
    assert np.array_equal(icp.Transform.identity().to_array(), oracle.transform_identity())
    assert np.allclose((T * T.inverse()).to_array(), oracle.transform_identity(), atol=1e-15)


def test_synthetic_generator_is_split_invariant():
    """Pair p of the synthetic stream is the same numbers whatever batch it is generated in —
    what lets every rank generate only its own slice (bench.py, shard.py)."""
    a_src, a_dst, off, a_T = icp.synth_scan_pairs(6, beams=90, first_pair=10, nthreads=3)
    b_src, b_dst, _, b_T = icp.synth_scan_pairs(2, beams=90, first_pair=13, nthreads=1)
    assert np.array_equal(a_src[3 * 90:5 * 90], b_src) and np.array_equal(a_dst[3 * 90:5 * 90], b_dst)
    assert np.array_equal(a_T[3:5], b_T)
    assert off[-1] == 6 * 90
    zero = (a_src == 0).all(axis=1).mean()
    assert 0.0 < zero < 0.05            # ~1 % invalid returns at the origin, like the bundled scans
