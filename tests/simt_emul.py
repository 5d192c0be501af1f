"""ctypes binding of the lock-step CTA emulator build of the kernel headers
(tests/simt/_build/libicp_simt.so).  CPU-only TEST INFRASTRUCTURE; see tests/simt/cpu_simt.h."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt")
_SO = os.path.join(_HERE, "_build", "libicp_simt.so")
_lib = None
_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
        _lib = C.CDLL(_SO)
    return _lib


def set_schedule(mode):
    lib().simt_set_schedule(C.c_int(mode))


_cost_buf = None


def set_options(guard_scale=0.0, want_cost=0, stream_ordered=False, shared_src=False, delta_band_scale=1.0):
    """Per-call state of the host library for the next icp_fused / icp_stream calls: knife-edge guard
    (bands x guard_scale, flagged pairs re-run with ordered sums), a final-cost output of `want_cost`
    pairs (read it back with last_cost()), reference-order sums in the streaming kernel,
    ICP_B200_FLAG_SHARED_SOURCE."""
    global _cost_buf
    _cost_buf = np.full(want_cost, np.nan) if want_cost else None
    lib().simt_set_options(C.c_double(guard_scale), _cost_buf.ctypes.data_as(_dp) if want_cost else None,
                           C.c_int(1 if stream_ordered else 0), C.c_int(1 if shared_src else 0),
                           C.c_double(delta_band_scale))


def set_fixed_point(on):
    """Outer-loop fixed-point exit of the kernels (icp_fused.cuh fused_repeats): on by default."""
    lib().simt_set_fixed_point(C.c_int(1 if on else 0))


def set_stream_nn_f32(on):
    """Streaming kernel's 3-D searches: binary32 filter on 16-byte records first (True, default) or
    every candidate in binary64."""
    lib().simt_set_stream_nn_f32(C.c_int(1 if on else 0))


def set_stream_nn_group(on):
    """Streaming kernel's search: a group of lanes per query (True) or one thread per query."""
    lib().simt_set_stream_nn_group(C.c_int(1 if on else 0))


def last_cost():
    return _cost_buf


def last_redo_count():
    lib().simt_last_redo_count.restype = C.c_uint32
    return int(lib().simt_last_redo_count())


def calc_stddevs(sets, mode=0, threads=128):
    """sets: list of (n,2) arrays -> (sigma (B,2), median (B,2)); sigma < 0 means None"""
    off = np.zeros(len(sets) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in sets])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64).reshape(-1, 2) for s in sets]
                                               + [np.zeros((1, 2))]))
    sig = np.zeros((len(sets), 2))
    med = np.zeros((len(sets), 2))
    lib().simt_calc_stddevs(flat.ctypes.data_as(_dp), off.ctypes.data_as(_u32p), C.c_uint32(len(sets)),
                            C.c_int(mode), C.c_int(threads), sig.ctypes.data_as(_dp), med.ctypes.data_as(_dp))
    return sig, med


def icp_fused(srcs, dsts, init_T, max_iter=20, variant=0, grid=1, corr_mode=2, cell_budget=4096, ordered=True):
    """srcs/dsts: lists of (n,dim) arrays.  Returns dict like oracle.icp_estimate_batch."""
    B = len(srcs)
    dim = srcs[0].shape[1]
    so = np.zeros(B + 1, dtype=np.uint32)
    do = np.zeros(B + 1, dtype=np.uint32)
    so[1:] = np.cumsum([len(s) for s in srcs])
    do[1:] = np.cumsum([len(d) for d in dsts])
    pad = np.zeros((1, dim))
    src = np.ascontiguousarray(np.concatenate(list(srcs) + [pad]), dtype=np.float64)
    dst = np.ascontiguousarray(np.concatenate(list(dsts) + [pad]), dtype=np.float64)
    T0 = np.ascontiguousarray(np.asarray(init_T, dtype=np.float64).reshape(B, 6))
    out_T = np.zeros((B, 6))
    inner = np.zeros((B, max(max_iter, 1)), dtype=np.uint32)
    ncorr = int(so[-1]) * (max_iter if corr_mode == 2 else 1)
    corr = np.full(max(ncorr, 1), -1, dtype=np.int32)
    hashes = np.zeros(B, dtype=np.uint64)
    status = np.zeros(B, dtype=np.int32)
    err = lib().simt_icp_fused(
        C.c_int(dim), C.c_int(variant), C.c_uint(grid), src.ctypes.data_as(_dp), so.ctypes.data_as(_u32p),
        dst.ctypes.data_as(_dp), do.ctypes.data_as(_u32p), T0.ctypes.data_as(_dp), C.c_uint32(B),
        C.c_uint32(max_iter), C.c_int(corr_mode), C.c_int(cell_budget), C.c_int(3 if ordered == 'svc' else 2 if ordered == 'lean' else (1 if ordered else 0)),
        out_T.ctypes.data_as(_dp),
        inner.ctypes.data_as(_u32p), corr.ctypes.data_as(C.POINTER(C.c_int32)),
        hashes.ctypes.data_as(C.POINTER(C.c_uint64)), status.ctypes.data_as(C.POINTER(C.c_int32)))
    return {"T": out_T, "inner": inner[:, :max_iter], "corr": corr, "corr_hash": hashes, "status": status,
            "err": err, "src_off": so}


def icp_stream(srcs, dsts, init_T, max_iter=20, threads=64, grid=1, corr_mode=2):
    """The streaming (large-scan) kernel, icp_stream.cuh.  Same result dict as icp_fused."""
    B = len(srcs)
    dim = srcs[0].shape[1]
    so = np.zeros(B + 1, dtype=np.uint32)
    do = np.zeros(B + 1, dtype=np.uint32)
    so[1:] = np.cumsum([len(s) for s in srcs])
    do[1:] = np.cumsum([len(d) for d in dsts])
    pad = np.zeros((1, dim))
    src = np.ascontiguousarray(np.concatenate(list(srcs) + [pad]), dtype=np.float64)
    dst = np.ascontiguousarray(np.concatenate(list(dsts) + [pad]), dtype=np.float64)
    T0 = np.ascontiguousarray(np.asarray(init_T, dtype=np.float64).reshape(B, 6))
    out_T = np.zeros((B, 6))
    inner = np.zeros((B, max(max_iter, 1)), dtype=np.uint32)
    ncorr = int(so[-1]) * (max_iter if corr_mode == 2 else 1)
    corr = np.full(max(ncorr, 1), -1, dtype=np.int32)
    hashes = np.zeros(B, dtype=np.uint64)
    status = np.zeros(B, dtype=np.int32)
    err = lib().simt_icp_stream(
        C.c_int(dim), C.c_int(threads), C.c_uint(grid), src.ctypes.data_as(_dp), so.ctypes.data_as(_u32p),
        dst.ctypes.data_as(_dp), do.ctypes.data_as(_u32p), T0.ctypes.data_as(_dp), C.c_uint32(B),
        C.c_uint32(max_iter), C.c_int(corr_mode), out_T.ctypes.data_as(_dp),
        inner.ctypes.data_as(_u32p), corr.ctypes.data_as(C.POINTER(C.c_int32)),
        hashes.ctypes.data_as(C.POINTER(C.c_uint64)), status.ctypes.data_as(C.POINTER(C.c_int32)))
    return {"T": out_T, "inner": inner[:, :max_iter], "corr": corr, "corr_hash": hashes, "status": status,
            "err": err, "src_off": so}


def target_search(dst, queries, budget, fast=True):
    """icp_target.cuh build kernels + nn_search_kernel on the emulator -> original indices (int32)."""
    d = np.ascontiguousarray(dst, dtype=np.float64)
    q = np.ascontiguousarray(queries, dtype=np.float64)
    out = np.full(q.shape[0], -1, dtype=np.int32)
    rc = lib().simt_target_search(C.c_int(d.shape[1]), d.ctypes.data_as(_dp), C.c_uint32(d.shape[0]),
                                  C.c_longlong(int(budget)), C.c_int({"warp": 2, "group4": 4, "group8": 8}.get(fast, 1 if fast else 0)), q.ctypes.data_as(_dp),
                                  C.c_uint64(q.shape[0]), out.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return out


def nn_counters(reset=True):
    """Work counters of the fast shared-target search since the last reset (emulator only)."""
    out = (C.c_ulonglong * 8)()
    lib().simt_nn_counters(out, C.c_int(1 if reset else 0))
    return dict(zip(("queries", "rows", "spans", "candidates", "pass2_queries", "exact_evals"), [int(v) for v in out]))


def grid_seeded_search(dst, queries, seeds, budget):
    """grid_nn_search_auto (icp_target_nn.cuh) on a cell list built by the target kernels.
    seeds: original index of any target point per query, or -1."""
    d = np.ascontiguousarray(dst, dtype=np.float64)
    q = np.ascontiguousarray(queries, dtype=np.float64)
    s = np.ascontiguousarray(seeds, dtype=np.int32)
    out = np.full(q.shape[0], -1, dtype=np.int32)
    rc = lib().simt_grid_seeded_search(C.c_int(d.shape[1]), d.ctypes.data_as(_dp), C.c_uint32(d.shape[0]),
                                       C.c_longlong(int(budget)), q.ctypes.data_as(_dp),
                                       s.ctypes.data_as(C.POINTER(C.c_int32)), C.c_uint64(q.shape[0]),
                                       out.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return out
