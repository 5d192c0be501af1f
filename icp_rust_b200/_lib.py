"""Loader of the native libraries.  The CUDA library is the product: if it is missing the package
fails loudly — there is no CPU fallback and no other backend."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("ICP_B200_LIB_DIR") or os.path.join(_HERE, "lib")   # override: debug builds (tools/)
CUDA_SO = os.path.join(LIB_DIR, "libicp_b200.so")
SYNTH_SO = os.path.join(LIB_DIR, "libicp_synth.so")


class MissingNativeLibrary(RuntimeError):
    pass


class Batch(C.Structure):
    """icp_b200_batch (include/icp_b200.h)"""
    _fields_ = [
        ("dim", C.c_uint32), ("batch", C.c_uint32), ("max_iter", C.c_uint32), ("corr_mode", C.c_int32),
        ("max_src_points", C.c_uint32), ("max_dst_points", C.c_uint32),
        ("src", C.c_void_p), ("src_off", C.c_void_p), ("dst", C.c_void_p), ("dst_off", C.c_void_p),
        ("init_T", C.c_void_p), ("out_T", C.c_void_p), ("out_inner", C.c_void_p), ("out_corr", C.c_void_p),
        ("out_corr_hash", C.c_void_p), ("out_status", C.c_void_p), ("flags", C.c_uint32),
        ("out_cost", C.c_void_p),
    ]


class CallInfo(C.Structure):
    """icp_b200_call_info"""
    _fields_ = [("kernel_launches", C.c_uint32), ("kernel_ms", C.c_float), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("grid_ctas", C.c_uint32), ("block_threads", C.c_uint32),
                ("smem_bytes", C.c_uint32), ("redo_pairs", C.c_uint32)]


# every symbol include/icp_b200.h declares
EXPORTS = [
    "icp_b200_abi_version", "icp_b200_status_string", "icp_b200_device_count", "icp_b200_create",
    "icp_b200_destroy", "icp_b200_last_call_info", "icp_b200_last_error", "icp_b200_estimate_batch",
    "icp_b200_estimate_batch_device", "icp_b200_icp2d_estimate", "icp_b200_icp3d_estimate",
    "icp_b200_target_create", "icp_b200_target_destroy", "icp_b200_target_estimate_batch",
    "icp_b200_target_estimate_batch_device", "icp_b200_target_search", "icp_b200_target_search_device",
    "icp_b200_calc_stddevs_batch", "icp_b200_target_update", "icp_b200_estimate_batch_submit",
    "icp_b200_estimate_batch_wait", "icp_b200_target_estimate_batch_submit",
]

_cuda = None
_synth = None


def cuda_lib():
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_SO):
            raise MissingNativeLibrary(
                f"{CUDA_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  icp_rust_b200 has no CPU fallback.")
        L = C.CDLL(CUDA_SO)
        L.icp_b200_status_string.restype = C.c_char_p
        L.icp_b200_status_string.argtypes = [C.c_int]
        L.icp_b200_last_error.restype = C.c_char_p
        L.icp_b200_last_error.argtypes = [C.c_void_p]
        L.icp_b200_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.icp_b200_destroy.argtypes = [C.c_void_p]
        L.icp_b200_destroy.restype = None
        L.icp_b200_last_call_info.argtypes = [C.c_void_p, C.POINTER(CallInfo)]
        L.icp_b200_estimate_batch.argtypes = [C.c_void_p, C.POINTER(Batch)]
        L.icp_b200_estimate_batch_device.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_int]
        L.icp_b200_estimate_batch_submit.argtypes = [C.c_void_p, C.POINTER(Batch)]
        L.icp_b200_estimate_batch_wait.argtypes = [C.c_void_p]
        L.icp_b200_target_estimate_batch_submit.argtypes = [C.c_void_p, C.POINTER(Batch)]
        for f in (L.icp_b200_icp2d_estimate, L.icp_b200_icp3d_estimate):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                          C.c_void_p]
        L.icp_b200_target_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.icp_b200_target_destroy.argtypes = [C.c_void_p]
        L.icp_b200_target_destroy.restype = None
        L.icp_b200_target_update.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.icp_b200_target_estimate_batch.argtypes = [C.c_void_p, C.POINTER(Batch)]
        L.icp_b200_target_estimate_batch_device.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_int]
        L.icp_b200_target_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.icp_b200_target_search_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                                    C.c_int]
        L.icp_b200_calc_stddevs_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                                  C.c_void_p]
        L.icp_b200_device_count.argtypes = [C.POINTER(C.c_int)]
        _cuda = L
    return _cuda


def synth_lib():
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_SO):
            raise MissingNativeLibrary(f"{SYNTH_SO} is missing: run __graft_entry__.build()")
        L = C.CDLL(SYNTH_SO)
        L.icp_synth_scan_pairs.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_double,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.icp_synth_map_scene.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
        L.icp_synth_map_scans.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        _synth = L
    return _synth
