"""icp_rust_b200 — host-side mirror of tier4/icp_rust's estimation API over libicp_b200.so.

The reference's public surface for the ICP hot path is (src/lib.rs:23-25, 91-174):
    Transform, Icp2d::new(dst).estimate(src, initial_transform, max_iter),
    Icp3d::new(dst).estimate(src, initial_transform, max_iter)
This module keeps those names and argument meanings, and adds the batched entry the device needs
(thousands of independent scan pairs per call).  All compute happens in hand-written sm_100a CUDA
kernels behind the C ABI of include/icp_b200.h; importing the package without the built library,
or calling it without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from ._lib import Batch, CallInfo, MissingNativeLibrary, cuda_lib, synth_lib, EXPORTS  # noqa: F401

CORR_NONE, CORR_LAST, CORR_ALL = 0, 1, 2
FLAG_ORDERED_SUMS = 1  # ICP_B200_FLAG_ORDERED_SUMS: reference-order sums, bit-identical transforms
FLAG_SHARED_SOURCE = 2  # one source scan, `batch` initial transforms
FLAG_NO_GUARD = 4  # default mode without the knife-edge guard / ordered re-run
FLAG_INDEPENDENT = 8  # asynchronous device-pointer call that touches nothing the previous call writes: may overlap its tail
ABI_VERSION = 4
BREAK_CAP, BREAK_NONE, BREAK_DELTA, BREAK_ERROR = 0, 1, 2, 3
PAIR_OK, PAIR_EMPTY_TARGET, PAIR_TOO_LARGE, PAIR_NON_FINITE = 0, 1, 2, 3
ERR_EMPTY_TARGET, ERR_NON_FINITE, ERR_TOO_LARGE = 4, 5, 6


class IcpError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        msg = cuda_lib().icp_b200_status_string(status).decode()
        super().__init__(f"icp_b200 status {status} ({msg}) {detail}".strip())


class Transform:
    """`icp::Transform` (src/transform.rs:6-51): rotation kept as a full 2x2 matrix plus translation.
    Host-side value type; the device returns it as 6 doubles [r00 r01 r10 r11 tx ty]."""

    __slots__ = ("rot", "t")

    def __init__(self, rot, t):
        self.rot = np.asarray(rot, dtype=np.float64).reshape(2, 2).copy()
        self.t = np.asarray(t, dtype=np.float64).reshape(2).copy()

    @staticmethod
    def new(param):  # transform.rs:13-16 -> se2::calc_rt (se2.rs:21-41)
        # Host-side convenience only (building an initial guess): sin / cos come from the platform libm here,
        # while the device kernels and the oracle evaluate the fdlibm restatement of icp_math.cuh — the two can
        # differ in the last bit.  Nothing on the compute path calls this; transforms the device returns are
        # passed around as their 6 doubles (from_array / to_array), never re-derived from parameters.
        vx, vy, theta = (float(v) for v in param)
        c, s = math.cos(theta), math.sin(theta)
        if theta == 0.0:
            t = (vx, vy)
        else:
            t = ((s * vx - (1.0 - c) * vy) / theta, ((1.0 - c) * vx + s * vy) / theta)
        return Transform([[c, -s], [s, c]], t)

    @staticmethod
    def from_rt(rot, t):  # transform.rs:18-20
        return Transform(rot, t)

    @staticmethod
    def identity():  # transform.rs:34-39
        return Transform(np.eye(2), np.zeros(2))

    @staticmethod
    def from_array(a):
        a = np.asarray(a, dtype=np.float64).reshape(6)
        return Transform(a[:4].reshape(2, 2), a[4:])

    def to_array(self):
        return np.concatenate([self.rot.reshape(4), self.t])

    def transform(self, landmark):  # transform.rs:22-24
        return self.rot @ np.asarray(landmark, dtype=np.float64) + self.t

    def inverse(self):  # transform.rs:26-32
        inv = self.rot.T
        return Transform(inv, -(inv @ self.t))

    def __mul__(self, rhs):  # transform.rs:42-51
        return Transform(self.rot @ rhs.rot, self.rot @ rhs.t + self.t)

    def __repr__(self):
        return f"Transform(rot={self.rot.tolist()}, t={self.t.tolist()})"


def _as_T6(T):
    if isinstance(T, Transform):
        return T.to_array()
    return np.ascontiguousarray(T, dtype=np.float64).reshape(6)


class Context:
    """One CUDA device (icp_b200_create).  Not thread-safe; use one per thread."""

    def __init__(self, device: int = 0):
        self._L = cuda_lib()
        h = C.c_void_p()
        rc = self._L.icp_b200_create(int(device), C.byref(h))
        if rc != 0:
            raise IcpError(rc, f"(icp_b200_create(device={device}))")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.icp_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise IcpError(rc, self._L.icp_b200_last_error(self._h).decode())
        return rc

    def call_info(self) -> dict:
        info = CallInfo()
        self._L.icp_b200_last_call_info(self._h, C.byref(info))
        return {k: getattr(info, k) for k, _ in CallInfo._fields_}

    # ------------------------------------------------------------------ batched estimate (host arrays)
    def estimate_batch(self, src, src_off, dst, dst_off, init_T, max_iter=20, *, dim=None, corr_mode=CORR_NONE,
                       want_inner=True, want_hash=True, target=None, raise_on_pair_error=True, ordered=False,
                       want_cost=False, shared_source=False, guard=True):
        """B independent `IcpD::new(dst_b).estimate(src_b, init_b, max_iter)` calls.
        src/dst: concatenated (P, dim) float64; *_off: (B+1,) uint32.  Returns a dict.
        shared_source: `src` is ONE scan (src_off = [0, n]) aligned from every row of init_T.
        want_cost: also return the final huber_error per pair (icp_b200_batch.out_cost).
        guard=False: ICP_B200_FLAG_NO_GUARD."""
        src = np.ascontiguousarray(src, dtype=np.float64)
        if dim is None:
            dim = src.shape[1] if src.ndim == 2 else 2
        src_off = np.ascontiguousarray(src_off, dtype=np.uint32)
        init_T = np.ascontiguousarray(init_T, dtype=np.float64).reshape(-1, 6)
        B = init_T.shape[0] if shared_source else src_off.size - 1
        assert init_T.shape[0] == B and (not shared_source or src_off.size == 2)
        out_T = np.zeros((B, 6))
        inner = np.zeros((B, max(max_iter, 1)), dtype=np.uint32) if want_inner else None
        hashes = np.zeros(B, dtype=np.uint64) if want_hash else None
        status = np.zeros(B, dtype=np.int32)
        cost = np.zeros(B) if want_cost else None
        nsrc = int(src_off[-1]) * (B if shared_source else 1)
        corr = None
        if corr_mode == CORR_ALL:
            corr = np.full(max(nsrc * max_iter, 1), -1, dtype=np.int32)
        elif corr_mode == CORR_LAST:
            corr = np.full(max(nsrc, 1), -1, dtype=np.int32)
        b = Batch()
        b.dim, b.batch, b.max_iter, b.corr_mode = dim, B, max_iter, corr_mode
        b.flags = (FLAG_ORDERED_SUMS if ordered else 0) | (FLAG_SHARED_SOURCE if shared_source else 0) | \
                  (0 if guard else FLAG_NO_GUARD)
        b.src = src.ctypes.data if src.size else None
        b.src_off = src_off.ctypes.data
        b.out_cost = cost.ctypes.data if cost is not None else None
        if target is None:
            dst = np.ascontiguousarray(dst, dtype=np.float64)
            dst_off = np.ascontiguousarray(dst_off, dtype=np.uint32)
            b.dst = dst.ctypes.data if dst.size else None
            b.dst_off = dst_off.ctypes.data
        b.init_T = init_T.ctypes.data
        b.out_T = out_T.ctypes.data
        b.out_inner = inner.ctypes.data if inner is not None else None
        b.out_corr = corr.ctypes.data if corr is not None else None
        b.out_corr_hash = hashes.ctypes.data if hashes is not None else None
        b.out_status = status.ctypes.data
        if target is None:
            rc = self._L.icp_b200_estimate_batch(self._h, C.byref(b))
        else:
            rc = self._L.icp_b200_target_estimate_batch(target._h, C.byref(b))
        allow = () if raise_on_pair_error else (ERR_EMPTY_TARGET, ERR_NON_FINITE, ERR_TOO_LARGE)
        self._check(rc, allow)
        res = {"T": out_T, "status": status, "rc": rc}
        if inner is not None:
            inner = inner[:, :max_iter]
            res.update(inner=inner, accepted=inner & 0xFFFF, reason=inner >> 16)
        if hashes is not None:
            res["corr_hash"] = hashes
        if corr is not None:
            res["corr"] = corr
        if cost is not None:
            res["cost"] = cost
        res["redo_pairs"] = self.call_info()["redo_pairs"]
        return res

    # ------------------------------------------------------------------ device-resident batch
    def estimate_batch_device(self, ptrs: dict, B, max_iter, dim=2, corr_mode=CORR_NONE, max_src=0, max_dst=0,
                              stream=None, synchronize=True, target=None, ordered=False, flags=0):
        """ptrs: device addresses (ints) for src, src_off, dst, dst_off, init_T, out_T and optionally
        out_inner, out_corr, out_corr_hash, out_status.  Used with torch tensors' data_ptr()."""
        b = Batch()
        b.dim, b.batch, b.max_iter, b.corr_mode = dim, B, max_iter, corr_mode
        b.max_src_points, b.max_dst_points = max_src, max_dst
        b.flags = (FLAG_ORDERED_SUMS if ordered else 0) | flags
        for k in ("src", "src_off", "dst", "dst_off", "init_T", "out_T", "out_inner", "out_corr", "out_corr_hash",
                  "out_status", "out_cost"):
            v = ptrs.get(k)
            setattr(b, k, v if v else None)
        if target is None:
            rc = self._L.icp_b200_estimate_batch_device(self._h, C.byref(b), stream, 1 if synchronize else 0)
        else:
            rc = self._L.icp_b200_target_estimate_batch_device(target._h, C.byref(b), stream, 1 if synchronize else 0)
        return self._check(rc)

    def estimate_batch_host_ptrs(self, ptrs: dict, B, max_iter, dim=2, corr_mode=CORR_NONE, target=None, ordered=False,
                                 flags=0):
        """icp_b200_estimate_batch on caller-owned HOST buffers given as addresses (e.g. pinned torch
        tensors): same keys as estimate_batch_device.  Copies in, computes, copies results back and
        returns when they have landed."""
        b = Batch()
        b.dim, b.batch, b.max_iter, b.corr_mode = dim, B, max_iter, corr_mode
        b.flags = (FLAG_ORDERED_SUMS if ordered else 0) | flags
        for k in ("src", "src_off", "dst", "dst_off", "init_T", "out_T", "out_inner", "out_corr", "out_corr_hash",
                  "out_status", "out_cost"):
            v = ptrs.get(k)
            setattr(b, k, v if v else None)
        if target is None:
            rc = self._L.icp_b200_estimate_batch(self._h, C.byref(b))
        else:
            rc = self._L.icp_b200_target_estimate_batch(target._h, C.byref(b))
        return self._check(rc)

    def submit_host_ptrs(self, ptrs: dict, B, max_iter, dim=2, corr_mode=CORR_NONE, target=None, ordered=False,
                         flags=0):
        """icp_b200_estimate_batch_submit: the first half of estimate_batch_host_ptrs — everything is enqueued,
        nothing waited for.  The buffers (pinned) stay valid and untouched until wait().  One thread that
        alternates two contexts keeps the device busy across batch boundaries."""
        b = Batch()
        b.dim, b.batch, b.max_iter, b.corr_mode = dim, B, max_iter, corr_mode
        b.flags = (FLAG_ORDERED_SUMS if ordered else 0) | flags
        for k in ("src", "src_off", "dst", "dst_off", "init_T", "out_T", "out_inner", "out_corr", "out_corr_hash",
                  "out_status", "out_cost"):
            v = ptrs.get(k)
            setattr(b, k, v if v else None)
        if target is None:
            rc = self._L.icp_b200_estimate_batch_submit(self._h, C.byref(b))
        else:
            rc = self._L.icp_b200_target_estimate_batch_submit(target._h, C.byref(b))
        return self._check(rc)

    def wait(self, allow=()):
        """icp_b200_estimate_batch_wait: completes the submitted batch; returns / raises what the one-call
        form would."""
        return self._check(self._L.icp_b200_estimate_batch_wait(self._h), allow)

    # ------------------------------------------------------------------ stats hook
    def calc_stddevs_batch(self, sets):
        """stats::calc_stddevs (stats.rs:49-60) of each (n,2) residual set -> (sigma, median), (B,2) each;
        sigma < 0 marks None."""
        off = np.zeros(len(sets) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(s) for s in sets])
        flat = np.ascontiguousarray(np.concatenate(
            [np.asarray(s, dtype=np.float64).reshape(-1, 2) for s in sets] + [np.zeros((1, 2))]))
        sig = np.zeros((len(sets), 2))
        med = np.zeros((len(sets), 2))
        self._check(self._L.icp_b200_calc_stddevs_batch(self._h, flat.ctypes.data, off.ctypes.data, len(sets),
                                                        sig.ctypes.data, med.ctypes.data))
        return sig, med


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Target:
    """Device-resident search structure of one target cloud (icp_b200_target_create): the value
    `Icp2d::new(dst)` / `Icp3d::new(dst)` returns (src/lib.rs:97-102, 139-144)."""

    def __init__(self, dst, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        dst = np.ascontiguousarray(dst, dtype=np.float64)
        if dst.ndim != 2 or dst.shape[1] not in (2, 3):
            raise ValueError("dst must be (M, 2) or (M, 3)")
        self.dim = dst.shape[1]
        self.m = dst.shape[0]
        h = C.c_void_p()
        rc = self.ctx._L.icp_b200_target_create(self.ctx._h, self.dim, dst.ctypes.data if dst.size else None,
                                                self.m, C.byref(h))
        self.ctx._check(rc)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._L.icp_b200_target_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, add=None, remove_idx=None):
        """icp_b200_target_update: drop the points `remove_idx` names (survivors close up), append `add`."""
        add = np.zeros((0, self.dim)) if add is None else np.ascontiguousarray(add, dtype=np.float64).reshape(-1, self.dim)
        rem = np.zeros(0, dtype=np.uint32) if remove_idx is None else np.ascontiguousarray(remove_idx, dtype=np.uint32)
        rc = self.ctx._L.icp_b200_target_update(self._h, add.ctypes.data if add.size else None, add.shape[0],
                                                rem.ctypes.data if rem.size else None, rem.size)
        self.ctx._check(rc)
        self.m = self.m - len(np.unique(rem)) + add.shape[0]

    def search(self, queries):
        """KdTree::search(&q).0 for every row of queries (src/lib.rs:121,164)."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, self.dim)
        out = np.zeros(q.shape[0], dtype=np.int32)
        self.ctx._check(self.ctx._L.icp_b200_target_search(self._h, q.ctypes.data, q.shape[0], out.ctypes.data))
        return out

    def search_device(self, queries_ptr: int, n_queries: int, out_index_ptr: int, stream=None, synchronize=True):
        """Same search on DEVICE buffers (addresses), enqueued on `stream`."""
        self.ctx._check(self.ctx._L.icp_b200_target_search_device(self._h, queries_ptr, n_queries, out_index_ptr,
                                                                  stream, 1 if synchronize else 0))

    def estimate_batch(self, src, src_off, init_T, max_iter=20, **kw):
        return self.ctx.estimate_batch(src, src_off, None, None, init_T, max_iter, dim=self.dim, target=self, **kw)


class _IcpBase:
    DIM = 2
    # targets up to this size ride inside the fused kernel's shared memory; larger ones get a
    # device-resident Target built once in `new`
    _INLINE_MAX = 4096

    def __init__(self, dst, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.dst = np.ascontiguousarray(dst, dtype=np.float64).reshape(-1, self.DIM)
        self._target = Target(self.dst, self.ctx) if self.dst.shape[0] > self._INLINE_MAX else None

    @classmethod
    def new(cls, dst, ctx=None):
        return cls(dst, ctx)

    def estimate(self, src, initial_transform, max_iter: int, ordered: bool = False) -> Transform:
        """Estimates the transform that converts the `src` points to `dst` (src/lib.rs:104-130 / 146-173).
        ordered=False: guarded warp-tree sums (the reference's decisions, transform within 1e-9, lowest
        latency); ordered=True: reference-order sums, bit-identical transform."""
        src = np.ascontiguousarray(src, dtype=np.float64).reshape(-1, self.DIM)
        T0 = _as_T6(initial_transform)
        off = np.array([0, len(src)], dtype=np.uint32)
        if self._target is not None:
            res = self._target.estimate_batch(src, off, T0[None], max_iter, want_inner=False, want_hash=False,
                                              ordered=ordered)
            return Transform.from_array(res["T"][0])
        if ordered:
            doff = np.array([0, len(self.dst)], dtype=np.uint32)
            res = self.ctx.estimate_batch(src, off, self.dst, doff, T0[None], max_iter, dim=self.DIM, want_inner=False,
                                          want_hash=False, ordered=True)
            return Transform.from_array(res["T"][0])
        out = np.zeros(6)
        L = self.ctx._L
        fn = L.icp_b200_icp2d_estimate if self.DIM == 2 else L.icp_b200_icp3d_estimate
        rc = fn(self.ctx._h, self.dst.ctypes.data if self.dst.size else None, self.dst.shape[0],
                src.ctypes.data if src.size else None, src.shape[0], T0.ctypes.data, int(max_iter), out.ctypes.data)
        self.ctx._check(rc)
        return Transform.from_array(out)


class Icp2d(_IcpBase):
    """`icp::Icp2d` (src/lib.rs:91-131)."""
    DIM = 2


class Icp3d(_IcpBase):
    """`icp::Icp3d` (src/lib.rs:133-174): NN search in xyz, SE(2) solve on xy."""
    DIM = 3


# ---------------------------------------------------------------------- workload tooling (bench / tests)
def synth_scan_pairs(count, beams=1080, seed=20240807, first_pair=0, noise_mm=3.0, dropout=0.01, nthreads=None,
                     out_src=None, out_dst=None):
    """Synthetic 2-D scan pairs of BASELINE.json config 2 (see csrc/synth.c).  Returns
    (src (count*beams,2), dst, offsets (count+1,), true_T (count,6))."""
    L = synth_lib()
    if (count + 1) * beams >= 2 ** 32:
        raise ValueError("count * beams exceeds the 32-bit offset tables of a batch")
    src = out_src if out_src is not None else np.empty((count * beams, 2))
    dst = out_dst if out_dst is not None else np.empty((count * beams, 2))
    tt = np.empty((count, 6))
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    L.icp_synth_scan_pairs(seed, first_pair, count, beams, noise_mm, dropout, src.ctypes.data, dst.ctypes.data,
                           tt.ctypes.data, nthreads)
    off = (np.arange(count + 1, dtype=np.uint64) * beams).astype(np.uint32)
    return src, dst, off, tt


def synth_map(map_points=100_000, seed=7):
    """Config 3 target: `map_points` world-frame points of one 40 m x 40 m scene."""
    L = synth_lib()
    m = np.empty((map_points, 2))
    L.icp_synth_map_scene(seed, map_points, m.ctypes.data)
    return m


def synth_map_scans(count, beams=1080, seed=7, nthreads=None):
    """Config 3 queries: `count` sensor-frame scans of the same scene + their perturbed poses."""
    L = synth_lib()
    scans = np.empty((count * beams, 2))
    T0 = np.empty((count, 6))
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    L.icp_synth_map_scans(seed, count, beams, scans.ctypes.data, T0.ctypes.data, nthreads)
    off = (np.arange(count + 1, dtype=np.uint64) * beams).astype(np.uint32)
    return scans, off, T0
