"""Pins the CPU oracle (oracle/icp_oracle.c) against every known-answer test the reference
holds for the ICP hot path.  Each test names the reference test it replays."""
import math

import numpy as np
import pytest

import oracle
from tests import refdata as R


# ---------------------------------------------------------------- src/stats.rs
@pytest.mark.parametrize("values,expected", R.MEDIAN_CASES)
def test_mutable_median(values, expected):  # stats.rs:69-90
    assert oracle.median(values) == expected


@pytest.mark.parametrize("values,expected", R.MAD_CASES)
def test_mutable_mad(values, expected):  # stats.rs:93-102
    assert oracle.mad(values) == expected


def test_mutable_standard_deviation():  # stats.rs:105-136
    s = oracle.standard_deviation(R.NORMAL_100)
    assert abs(s - R.NORMAL_100_NUMPY_STD) < 0.5


def test_calc_stddevs():  # stats.rs:139-180
    s = oracle.calc_stddevs(R.MEASUREMENTS_30)
    assert abs(s[0] - R.MEASUREMENTS_30_NUMPY_STD[0]) < 1.0
    assert abs(s[1] - R.MEASUREMENTS_30_NUMPY_STD[1]) < 1.0


def test_median_is_the_order_statistic_mean():
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 10, 11, 665, 666, 1080, 1081):
        v = rng.normal(0, 3, n)
        s = np.sort(v)
        exp = s[n // 2] if n % 2 else (s[n // 2 - 1] + s[n // 2]) / 2.0
        assert oracle.median(v) == exp
        a = np.sort(np.abs(v - exp))
        mad = a[n // 2] if n % 2 else (a[n // 2 - 1] + a[n // 2]) / 2.0
        assert oracle.mad(v) == mad
        assert oracle.standard_deviation(v) == 1.482602218505602 * mad


# ---------------------------------------------------------------- src/huber.rs
def test_rho():  # huber.rs:33-37
    assert oracle.huber_rho(0.1 * 0.1, 0.1) == 0.1 * 0.1
    assert oracle.huber_rho(0.101 * 0.101, 0.1) == 2. * 0.1 * 0.101 - 0.1 * 0.1
    assert oracle.huber_rho(0.09 * 0.09, 0.1) == 0.09 * 0.09


def test_drho():  # huber.rs:40-70
    rho, drho = oracle.huber_rho, oracle.huber_drho

    def fd(a, h, k):
        e1, e0 = (a + h) ** 2, a ** 2
        return (rho(e1, k) - rho(e0, k)) / (e1 - e0), e0

    exp, e0 = fd(4.000, 0.001, 4.0)
    assert abs(drho(e0, 4.0) - exp) < 1e-3
    exp, e0 = fd(0.10, 0.01, 4.0)
    assert exp == drho(e0, 4.0)
    exp, e0 = fd(0.10, 0.0001, 0.10)
    assert abs(drho(e0, 0.10) - exp) < 1e-3
    exp, e0 = fd(5.000, 0.001, 4.0)
    assert abs(drho(e0, 4.0) - exp) < 1e-3
    exp, e0 = fd(10.000, 0.001, 4.0)
    assert abs(drho(e0, 4.0) - exp) < 1e-3


# ---------------------------------------------------------------- src/linalg.rs
def test_inverse3x3():  # linalg.rs:37-72
    inv = oracle.inverse3x3(R.INV_A)
    assert oracle.norm(inv @ R.INV_A - np.eye(3)) < 1e-14
    assert oracle.inverse3x3(np.zeros((3, 3))) is None
    assert oracle.inverse3x3(R.INV_RANK2) is None
    inv = oracle.inverse3x3(R.INV_TINY)
    assert inv is not None
    assert oracle.norm(inv @ R.INV_TINY - np.eye(3)) < 1e-14


# ---------------------------------------------------------------- src/so2.rs
def test_so2_exp():  # so2.rs:39-48
    rot = oracle.so2_exp(0.3)
    assert rot.shape == (2, 2)
    assert rot[0, 0] == oracle.cos(0.3) and rot[0, 1] == -oracle.sin(0.3)
    assert rot[1, 0] == oracle.sin(0.3) and rot[1, 1] == oracle.cos(0.3)
    # and the restated libm agrees with the platform libm to the last ulp or better
    assert abs(rot[0, 0] - math.cos(0.3)) <= np.spacing(math.cos(0.3))
    assert abs(rot[1, 0] - math.sin(0.3)) <= np.spacing(math.sin(0.3))


def test_so2_log():  # so2.rs:51-68
    for f in (0.3, 0.8, -0.7, -0.1):
        theta = f * math.pi
        assert abs(oracle.so2_log(oracle.so2_exp(theta)) - theta) < 1e-6


# ---------------------------------------------------------------- src/se2.rs
@pytest.mark.parametrize("param,expected", R.SE2_EXP_CASES)
def test_se2_exp(param, expected):  # se2.rs:85-142
    assert oracle.norm(oracle.se2_exp(param) - np.array(expected)) < 1e-6


@pytest.mark.parametrize("mat,expected", R.SE2_LOG_CASES)
def test_se2_log(mat, expected):  # se2.rs:145-200
    assert oracle.norm(oracle.se2_log(mat) - np.array(expected)) < 1e-6


def test_se2_get_rt():  # se2.rs:203-220
    m = np.array(R.SE2_EXP_CASES[0][1])
    rot, t = oracle.se2_get_rt(m)
    assert np.array_equal(rot, m[:2, :2])
    assert np.array_equal(t, m[:2, 2])


def test_sincos_matches_libm_within_one_ulp():
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-0.8, 0.8, 3000), rng.uniform(-20, 20, 3000),
                         rng.uniform(-1e5, 1e5, 2000),
                         np.array([k * math.pi / 2 + e for k in range(-8, 9)
                                   for e in (-1e-7, 0.0, 1e-7)])])
    for x in xs:
        for mine, ref in ((oracle.sin(x), math.sin(x)), (oracle.cos(x), math.cos(x))):
            assert abs(mine - ref) <= np.spacing(abs(ref)), x
    assert oracle.sin(0.0) == 0.0 and oracle.cos(0.0) == 1.0
    assert math.copysign(1.0, oracle.sin(-0.0)) == -1.0
    assert math.isnan(oracle.sin(float("inf"))) and math.isnan(oracle.cos(float("nan")))


# ---------------------------------------------------------------- src/transform.rs
def test_transform():  # transform.rs:62-70
    T = oracle.transform_from_rt(oracle.so2_new_rotation2(math.pi / 2), [3., 6.])
    assert oracle.norm(oracle.transform_apply(T, [4., 2.]) - np.array([-2. + 3., 4. + 6.])) < 1e-8


def test_transform_inverse():  # transform.rs:73-80
    T = oracle.transform_inverse(
        oracle.transform_from_rt(oracle.so2_new_rotation2(math.pi / 2), [3., 6.]))
    assert oracle.norm(oracle.transform_apply(T, [-2. + 3., 4. + 6.]) - np.array([4., 2.])) < 1e-8


def test_transform_mul():  # transform.rs:83-96
    T1 = oracle.transform_inverse(
        oracle.transform_from_rt(oracle.so2_new_rotation2(math.pi / 4), [2., 1.]))
    T2 = oracle.transform_inverse(
        oracle.transform_from_rt(oracle.so2_new_rotation2(math.pi / 2), [5., 3.]))
    x = [-5., 6.]
    pa = oracle.transform_apply(T1, oracle.transform_apply(T2, x))
    pb = oracle.transform_apply(oracle.transform_mul(T1, T2), x)
    assert oracle.norm(pa - pb) < 1e-8


# ---------------------------------------------------------------- src/lib.rs
def test_residual():  # lib.rs:267-274
    T = oracle.transform_new([-10., 20., 0.01])
    src = [7., 8.]
    dst = oracle.transform_apply(T, src)
    assert np.array_equal(oracle.residual(T, src, dst), np.zeros(2))


def test_error():  # lib.rs:277-297
    src = np.array([[-6., 9.], [-1., 9.], [-4., -4.]])
    dst = np.array([[-4., 4.], [0., 3.], [-3., -8.]])
    T = oracle.transform_new([10., 20., 0.01])
    r = [oracle.residual(T, s, d) for s, d in zip(src, dst)]
    dot = lambda v: float(v[0]) * float(v[0]) + float(v[1]) * float(v[1])  # no FMA, unlike BLAS
    expected = (dot(r[0]) + dot(r[1])) + dot(r[2])
    assert oracle.error(T, src, dst) == expected


def _apply_all(T, pts):
    return np.array([oracle.transform_apply(T, p) for p in pts]).reshape(-1, 2)


def test_gauss_newton_update_input_size():  # lib.rs:300-318
    T = oracle.transform_new([10.0, 30.0, -0.15])
    assert oracle.gauss_newton_update(T, np.zeros((0, 2)), np.zeros((0, 2))) is None
    src = R.WGN_SIZE_SRCS[1]
    assert oracle.gauss_newton_update(T, src, _apply_all(T, src)) is None
    src = R.WGN_SIZE_SRCS[2]
    assert oracle.gauss_newton_update(T, src, _apply_all(T, src)) is not None


def test_gauss_newton_update():  # lib.rs:321-351
    true_param = np.array([10.0, 30.0, -0.15])
    initial_param = true_param + np.array([0.3, -0.5, 0.001])
    Ttrue = oracle.transform_new(true_param)
    Tinit = oracle.transform_new(initial_param)
    dst = _apply_all(Ttrue, R.GN_SRC5)
    upd = oracle.gauss_newton_update(Tinit, R.GN_SRC5, dst)
    assert upd is not None
    Tupd = oracle.transform_new(initial_param + upd)
    assert oracle.error(Tupd, R.GN_SRC5, dst) < oracle.error(Tinit, R.GN_SRC5, dst) * 0.01


def test_weighted_gauss_newton_update_input_size():  # lib.rs:354-401
    T = oracle.transform_new(R.WGN_SIZE_PARAM)
    for src in R.WGN_SIZE_SRCS:
        dst = _apply_all(T, src)
        assert oracle.weighted_gauss_newton_update(T, src, dst) is None


def test_weighted_gauss_newton_update_zero_x_diff():  # lib.rs:404-427
    Ttrue = oracle.transform_new([0.00, 0.01, 0.00])
    dst = _apply_all(Ttrue, R.ZERO_X_SRC)
    Tinit = oracle.transform_new([0.00, 0.00, 0.00])
    assert oracle.weighted_gauss_newton_update(Tinit, R.ZERO_X_SRC, dst) is None


def test_weighted_gauss_newton_update():  # lib.rs:430-507
    true_param = np.array([10.0, 30.0, -0.15])
    initial_param = true_param + np.array([0.3, -0.5, 0.001])
    Ttrue = oracle.transform_new(true_param)
    Tinit = oracle.transform_new(initial_param)
    src = R.WGN_SRC19
    dst = _apply_all(Ttrue, src) + R.WGN_NOISE19
    upd = oracle.weighted_gauss_newton_update(Tinit, src, dst)
    assert upd is not None
    Tupd = oracle.transform_new(initial_param + upd)
    e0 = oracle.error(Tinit, src, dst)
    assert oracle.error(Tupd, src, dst) < e0 * 0.1
    Test = oracle.estimate_transform(src, dst)
    assert oracle.error(Test, src, dst) < e0 * 0.001


@pytest.mark.parametrize("nn_mode", [oracle.NN_BRUTE, oracle.NN_KDTREE])
def test_icp_2dscan(nn_mode):  # lib.rs:554-595
    src = R.L_SHAPE_2D
    Ttrue = oracle.transform_new(R.L_TRUE_PARAM)
    dst = _apply_all(Ttrue, src)
    Tinit = oracle.transform_mul(oracle.transform_new(R.L_NOISE_PARAM), Ttrue)
    res = oracle.icp_estimate(src, dst, Tinit, 20, nn_mode=nn_mode)
    assert res["status"] == 0
    for sp, dp in zip(src, dst):
        assert oracle.norm(oracle.transform_apply(res["T"], sp) - dp) < 1e-3


@pytest.mark.parametrize("nn_mode", [oracle.NN_BRUTE, oracle.NN_KDTREE])
def test_icp_3dscan(nn_mode):  # lib.rs:510-551
    src = R.L_SHAPE_3D
    Ttrue = oracle.transform_new(R.L_TRUE_PARAM)
    dst = np.array([oracle.transform_xy(Ttrue, p) for p in src])
    Tinit = oracle.transform_mul(oracle.transform_new(R.L_NOISE_PARAM), Ttrue)
    res = oracle.icp_estimate(src, dst, Tinit, 20, nn_mode=nn_mode)
    for sp, dp in zip(src, dst):
        assert oracle.norm(oracle.transform_xy(res["T"], sp) - dp) < 1e-3


# ---------------------------------------------------------------- NN contract + bundled scans
def test_kdtree_equals_brute_force_with_ties():
    rng = np.random.default_rng(3)
    for dim in (2, 3):
        dst = rng.integers(-4, 5, size=(300, dim)).astype(np.float64)  # many exact ties
        q = rng.integers(-6, 7, size=(500, dim)).astype(np.float64) * 0.5
        a = oracle.nn_search(dst, q, oracle.NN_BRUTE)
        b = oracle.nn_search(dst, q, oracle.NN_KDTREE)
        assert np.array_equal(a, b)
        d2 = ((q[:, None, :] - dst[None, :, :]) ** 2).sum(-1)
        assert np.array_equal(a, d2.argmin(1))  # numpy argmin = lowest index on ties


def test_bundled_scan_chain_is_self_consistent(scans2d):
    """examples/scan2d.rs:62-88 call pattern on the first frames: kd-tree and brute-force
    oracles agree bit for bit (indices via hash, iteration trace, transform)."""
    src = scans2d[1]
    T = oracle.transform_identity()
    for k in (2, 3, 4):
        a = oracle.icp_estimate(src, scans2d[k], T, 20, nn_mode=oracle.NN_BRUTE)
        b = oracle.icp_estimate(src, scans2d[k], T, 20, nn_mode=oracle.NN_KDTREE)
        assert a["corr_hash"] == b["corr_hash"]
        assert np.array_equal(a["inner"], b["inner"])
        assert np.array_equal(a["T"], b["T"])
        T = a["T"]
    # frame-to-frame motion of the bundled data is a few millimetres / milliradians
    assert abs(T[4]) < 100 and abs(T[5]) < 100 and abs(T[2]) < 0.05


def test_duplicate_origin_points_pick_lowest_index(scans2d):
    dst = scans2d[99]
    zeros = np.where((dst == 0).all(axis=1))[0]
    assert len(zeros) >= 2  # scans >= 099 hold duplicated (0,0) returns
    q = np.array([[0.1, -0.2], [0.0, 0.0]])
    for mode in (oracle.NN_BRUTE, oracle.NN_KDTREE):
        idx = oracle.nn_search(dst, q, mode)
        assert idx[1] == zeros[0]


def test_empty_inputs():
    T0 = oracle.transform_new([1.0, 2.0, 0.3])
    r = oracle.icp_estimate(np.zeros((0, 2)), np.zeros((0, 2)), T0, 5)
    assert r["status"] == 0 and np.array_equal(r["T"], T0)
    r = oracle.icp_estimate(np.zeros((3, 2)), np.zeros((0, 2)), T0, 5)
    assert r["status"] == 1  # reference panics on unwrap (lib.rs:122)
    # a single source point: check_input_size fails -> identity update every iteration
    r = oracle.icp_estimate(np.array([[1.0, 1.0]]), np.array([[1.5, 1.0], [3.0, 3.0]]), T0, 5)
    assert np.array_equal(r["T"], T0) and (r["reason"] == oracle.BREAK_NONE).all()
