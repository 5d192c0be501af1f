"""ctypes binding of the CPU ORACLE (oracle/icp_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (icp_rust_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libicp_oracle.so")

BREAK_CAP, BREAK_NONE, BREAK_DELTA, BREAK_ERROR = 0, 1, 2, 3
NN_BRUTE, NN_KDTREE = 0, 1
CORR_NONE, CORR_LAST, CORR_ALL = 0, 1, 2
HUBER_K = 1.345


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("icp_oracle.c", "icp_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_sin.restype = C.c_double
        L.orc_sin.argtypes = [C.c_double]
        L.orc_cos.restype = C.c_double
        L.orc_cos.argtypes = [C.c_double]
        L.orc_huber_rho.restype = C.c_double
        L.orc_huber_rho.argtypes = [C.c_double, C.c_double]
        L.orc_huber_drho.restype = C.c_double
        L.orc_huber_drho.argtypes = [C.c_double, C.c_double]
        L.orc_so2_log.restype = C.c_double
        L.orc_norm.restype = C.c_double
        L.orc_error.restype = C.c_double
        L.orc_huber_error.restype = C.c_double
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64]
        L.orc_nn_brute.restype = C.c_int64
        L.orc_kdtree_new.restype = C.c_void_p
        L.orc_kdtree_search.restype = C.c_int64
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint32))


# ------------------------------------------------------------------ scalar maths
def sin(x: float) -> float:
    return lib().orc_sin(float(x))


def cos(x: float) -> float:
    return lib().orc_cos(float(x))


def huber_rho(e, k=HUBER_K):
    return lib().orc_huber_rho(float(e), float(k))


def huber_drho(e, k=HUBER_K):
    return lib().orc_huber_drho(float(e), float(k))


# ------------------------------------------------------------------ stats.rs
def _opt1(fn, values):
    buf, p = _d(np.array(values, dtype=np.float64).copy())
    out = C.c_double()
    ok = fn(p, C.c_size_t(buf.size), C.byref(out))
    return out.value if ok else None


def median(values):
    return _opt1(lib().orc_median, values)


def mad(values):
    return _opt1(lib().orc_mad, values)


def standard_deviation(values):
    return _opt1(lib().orc_standard_deviation, values)


def calc_stddevs(residuals):
    r, p = _d(residuals)
    n, dim = r.shape
    out = np.zeros(dim)
    ok = lib().orc_calc_stddevs(p, C.c_size_t(n), C.c_int(dim), out.ctypes.data_as(_dp))
    return out if ok else None


# ------------------------------------------------------------------ linalg / so2 / se2
def inverse3x3(m):
    a, p = _d(np.asarray(m).reshape(9))
    out = np.zeros(9)
    ok = lib().orc_inverse3x3(p, out.ctypes.data_as(_dp))
    return out.reshape(3, 3) if ok else None


def so2_new_rotation2(theta):
    out = np.zeros(4)
    lib().orc_so2_new_rotation2(C.c_double(theta), out.ctypes.data_as(_dp))
    return out.reshape(2, 2)


def so2_exp(theta):
    out = np.zeros(4)
    lib().orc_so2_exp(C.c_double(theta), out.ctypes.data_as(_dp))
    return out.reshape(2, 2)


def so2_log(rot):
    a, p = _d(np.asarray(rot).reshape(4))
    return lib().orc_so2_log(p)


def se2_calc_rt(param):
    a, p = _d(param)
    rot, t = np.zeros(4), np.zeros(2)
    lib().orc_se2_calc_rt(p, rot.ctypes.data_as(_dp), t.ctypes.data_as(_dp))
    return rot.reshape(2, 2), t


def se2_exp(param):
    a, p = _d(param)
    out = np.zeros(9)
    lib().orc_se2_exp(p, out.ctypes.data_as(_dp))
    return out.reshape(3, 3)


def se2_log(m):
    a, p = _d(np.asarray(m).reshape(9))
    out = np.zeros(3)
    lib().orc_se2_log(p, out.ctypes.data_as(_dp))
    return out


def se2_get_rt(m):
    a, p = _d(np.asarray(m).reshape(9))
    rot, t = np.zeros(4), np.zeros(2)
    lib().orc_se2_get_rt(p, rot.ctypes.data_as(_dp), t.ctypes.data_as(_dp))
    return rot.reshape(2, 2), t


# ------------------------------------------------------------------ transform.rs  (T = 6 doubles)
def transform_new(param):
    a, p = _d(param)
    T = np.zeros(6)
    lib().orc_transform_new(p, T.ctypes.data_as(_dp))
    return T


def transform_identity():
    T = np.zeros(6)
    lib().orc_transform_identity(T.ctypes.data_as(_dp))
    return T


def transform_from_rt(rot, t):
    return np.concatenate([np.asarray(rot, dtype=np.float64).reshape(4),
                           np.asarray(t, dtype=np.float64).reshape(2)])


def transform_apply(T, pnt):
    a, pa = _d(T)
    b, pb = _d(pnt)
    out = np.zeros(2)
    lib().orc_transform_apply(pa, pb, out.ctypes.data_as(_dp))
    return out


def transform_inverse(T):
    a, pa = _d(T)
    out = np.zeros(6)
    lib().orc_transform_inverse(pa, out.ctypes.data_as(_dp))
    return out


def transform_mul(A, B):
    a, pa = _d(A)
    b, pb = _d(B)
    out = np.zeros(6)
    lib().orc_transform_mul(pa, pb, out.ctypes.data_as(_dp))
    return out


def transform_xy(T, p3):
    """src/lib.rs:52-57"""
    xy = transform_apply(T, np.asarray(p3, dtype=np.float64)[:2])
    return np.array([xy[0], xy[1], p3[2]])


def norm(v):
    a, p = _d(np.asarray(v).reshape(-1))
    return lib().orc_norm(p, C.c_size_t(a.size))


# ------------------------------------------------------------------ lib.rs solver
def residual(T, s, d):
    a, pa = _d(T)
    b, pb = _d(s)
    c, pc = _d(d)
    out = np.zeros(2)
    lib().orc_residual(pa, pb, pc, out.ctypes.data_as(_dp))
    return out


def error(T, src, dst):
    a, pa = _d(T)
    s, ps = _d(src)
    d, pd = _d(dst)
    return lib().orc_error(pa, ps, pd, C.c_size_t(s.size // 2))


def huber_error(T, src, dst):
    a, pa = _d(T)
    s, ps = _d(src)
    d, pd = _d(dst)
    return lib().orc_huber_error(pa, ps, pd, C.c_size_t(s.size // 2))


def _update(fn, T, src, dst):
    a, pa = _d(T)
    s, ps = _d(np.asarray(src, dtype=np.float64).reshape(-1))
    d, pd = _d(np.asarray(dst, dtype=np.float64).reshape(-1))
    assert s.size == d.size
    delta = np.zeros(3)
    ok = fn(pa, ps, pd, C.c_size_t(s.size // 2), delta.ctypes.data_as(_dp))
    return delta if ok else None


def gauss_newton_update(T, src, dst):
    return _update(lib().orc_gauss_newton_update, T, src, dst)


def weighted_gauss_newton_update(T, src, dst):
    return _update(lib().orc_weighted_gauss_newton_update, T, src, dst)


def estimate_transform(src, dst, trace=False):
    s, ps = _d(np.asarray(src, dtype=np.float64).reshape(-1))
    d, pd = _d(np.asarray(dst, dtype=np.float64).reshape(-1))
    T = np.zeros(6)
    acc, why = C.c_uint32(), C.c_uint32()
    lib().orc_estimate_transform(ps, pd, C.c_size_t(s.size // 2), T.ctypes.data_as(_dp),
                                 C.byref(acc), C.byref(why))
    return (T, acc.value, why.value) if trace else T


# ------------------------------------------------------------------ nearest neighbour
def nn_search(dst, queries, mode=NN_BRUTE):
    d, pd = _d(dst)
    q, pq = _d(queries)
    m, dim = d.shape
    out = np.zeros(q.shape[0], dtype=np.int32)
    lib().orc_nn_search_batch(pd, C.c_size_t(m), C.c_int(dim), pq, C.c_size_t(q.shape[0]),
                              C.c_int(mode), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


# ------------------------------------------------------------------ ICP drivers
class _Trace(C.Structure):
    _fields_ = [("nn_mode", C.c_int), ("corr_mode", C.c_int),
                ("corr", C.POINTER(C.c_int32)), ("inner", C.POINTER(C.c_uint32)),
                ("corr_hash", C.POINTER(C.c_uint64)), ("shared_tree", C.c_void_p),
                ("final_cost", C.POINTER(C.c_double)), ("margins", C.POINTER(C.c_double))]


def icp_estimate(src, dst, init_T=None, max_iter=20, nn_mode=NN_BRUTE, corr_mode=CORR_NONE):
    """Icp2d::new(dst).estimate(src, init, max_iter) (src/lib.rs:97-130) or the Icp3d twin
    (src/lib.rs:139-173), chosen by the point dimension.  Returns a dict."""
    s, ps = _d(src)
    d, pd = _d(dst)
    dim = s.shape[1] if s.ndim == 2 else d.shape[1]
    n = s.shape[0] if s.ndim == 2 else 0
    m = d.shape[0] if d.ndim == 2 else 0
    T0 = transform_identity() if init_T is None else np.ascontiguousarray(init_T, dtype=np.float64)
    out = np.zeros(6)
    inner = np.zeros(max(max_iter, 1), dtype=np.uint32)
    ncorr = n * max_iter if corr_mode == CORR_ALL else n
    corr = np.full(max(ncorr, 1), -1, dtype=np.int32)
    h = C.c_uint64(0)
    cost = C.c_double(0.)
    tr = _Trace(nn_mode, corr_mode, corr.ctypes.data_as(C.POINTER(C.c_int32)),
                inner.ctypes.data_as(C.POINTER(C.c_uint32)), C.pointer(h), None, C.pointer(cost), None)
    fn = lib().orc_icp2d_estimate if dim == 2 else lib().orc_icp3d_estimate
    st = fn(ps, C.c_size_t(n), pd, C.c_size_t(m), T0.ctypes.data_as(_dp), C.c_uint32(max_iter),
            out.ctypes.data_as(_dp), C.byref(tr))
    inner = inner[:max_iter]
    res = {"T": out, "status": st, "accepted": inner & 0xFFFF, "reason": inner >> 16,
           "inner": inner, "corr_hash": h.value, "cost": cost.value}
    if corr_mode == CORR_ALL:
        res["corr"] = corr[:ncorr].reshape(max_iter, n)
    elif corr_mode == CORR_LAST:
        res["corr"] = corr[:n]
    return res


def icp_estimate_batch(dim, src, src_off, dst, dst_off, init_T, max_iter=20, nn_mode=NN_KDTREE,
                       nthreads=None, want_cost=False, want_margins=False):
    """Ragged batch over pthreads; returns dict(T, inner, corr_hash, status, threads).
    dst_off=None: every pair is estimated against the one shared target `dst` (kd-tree built once,
    as Icp2d::new(&dst) followed by many estimate calls).  want_cost: huber_error of the last inner
    solve (see orc_trace.final_cost).  want_margins: (B,2) smallest relative distance of a
    delta-norm / error break test from its threshold."""
    s, ps = _d(np.asarray(src).reshape(-1))
    d, pd = _d(np.asarray(dst).reshape(-1))
    so, pso = _u32(src_off)
    B = so.size - 1
    if dst_off is None:
        pdo, shared_m = None, d.size // dim
    else:
        do, pdo = _u32(dst_off)
        shared_m = 0
    T0, pT0 = _d(np.asarray(init_T).reshape(-1))
    assert T0.size == 6 * B
    out = np.zeros((B, 6))
    inner = np.zeros((B, max_iter), dtype=np.uint32)
    hashes = np.zeros(B, dtype=np.uint64)
    status = np.zeros(B, dtype=np.int32)
    cost = np.zeros(B) if want_cost else None
    margins = np.zeros((B, 2)) if want_margins else None
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    used = lib().orc_icp_estimate_batch_ex(
        C.c_int(dim), ps, pso, pd, pdo, C.c_size_t(shared_m), pT0, C.c_uint32(B), C.c_uint32(max_iter),
        C.c_int(nn_mode), C.c_int(nthreads), out.ctypes.data_as(_dp),
        inner.ctypes.data_as(C.POINTER(C.c_uint32)),
        hashes.ctypes.data_as(C.POINTER(C.c_uint64)),
        status.ctypes.data_as(C.POINTER(C.c_int32)),
        cost.ctypes.data_as(_dp) if cost is not None else None,
        margins.ctypes.data_as(_dp) if margins is not None else None)
    res = {"T": out, "inner": inner, "corr_hash": hashes, "status": status, "threads": used}
    if cost is not None:
        res["cost"] = cost
    if margins is not None:
        res["margins"] = margins
    return res
