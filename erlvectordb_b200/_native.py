"""ctypes binding of libevdb_b200.so (the C ABI declared in include/evdb.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is
usable, every store operation raises.  The library is loaded from this package
directory (built in-tree by ``erlvectordb_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# EVDB_LIB_PATH: A/B timing of two builds of the same ABI (tools/); never a fallback
LIB_PATH = os.environ.get("EVDB_LIB_PATH") or os.path.join(PKG, "libevdb_b200.so")

F32, BF16, U8, U4 = 0, 1, 2, 3
COSINE, EUCLIDEAN, MANHATTAN = 0, 1, 2
PLAN_AUTO, PLAN_SCAN, PLAN_GEMM, PLAN_EXACT = 0, 1, 2, 3
DTYPES = {"f32": F32, "bf16": BF16, "u8": U8, "u4": U4,
          "quantization_8bit": U8, "quantization_4bit": U4}
METRICS = {"cosine": COSINE, "euclidean": EUCLIDEAN, "manhattan": MANHATTAN}
PLANS = {"auto": PLAN_AUTO, "scan": PLAN_SCAN, "gemm": PLAN_GEMM, "exact": PLAN_EXACT}

OK = 0
E_DIM_MISMATCH, E_BAD_VECTOR, E_OOM, E_CUDA, E_NCCL = -1, -2, -3, -4, -5
E_BAD_ARG, E_NO_DEVICE, E_UNSUPPORTED, E_BADARITH = -6, -7, -8, -9


class EvdbError(RuntimeError):
    def __init__(self, code: int, where: str = ""):
        self.code = code
        L = lib()
        name = L.evdb_strerror(code).decode()
        detail = L.evdb_last_cuda_error().decode()
        super().__init__(f"{where}: {name} ({code})" + (f" [{detail}]" if detail and code in (E_CUDA, E_OOM) else ""))


class Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("dtype", C.c_int32), ("dim", C.c_int32),
                ("gemm_shadow", C.c_int32), ("capacity_hint", C.c_uint64),
                ("n_shards", C.c_int32), ("devices", C.c_int32 * 16)]


class SearchOpts(C.Structure):
    _fields_ = [("plan", C.c_int32), ("kp_min", C.c_int32), ("slot_base", C.c_uint64), ("slot_stride", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("count", C.c_uint64), ("dimension", C.c_int32), ("dtype", C.c_int32),
                ("device", C.c_int32), ("last_plan", C.c_int32), ("capacity", C.c_uint64),
                ("device_bytes", C.c_uint64), ("searches", C.c_uint64),
                ("rows_scanned", C.c_uint64), ("escalations", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("last_search_ms", C.c_double),
                ("n_shards", C.c_int32), ("gemm_disabled", C.c_int32), ("shadow_bytes", C.c_uint64),
                ("upserts", C.c_uint64), ("deletes", C.c_uint64), ("last_h2d_ms", C.c_double),
                ("last_device_ms", C.c_double), ("last_d2h_ms", C.c_double)]


# every symbol include/evdb.h declares: (name, restype, argtypes)
_vp, _i, _u32, _u64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64
_pd, _pf, _pu8 = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
_pu32, _pi32, _pi64 = C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
SYMBOLS = [
    ("evdb_abi_version", _i, []),
    ("evdb_init", _i, [C.POINTER(C.c_int), _i]),
    ("evdb_strerror", C.c_char_p, [_i]),
    ("evdb_last_cuda_error", C.c_char_p, []),
    ("evdb_store_create", _i, [C.POINTER(Opts), C.POINTER(_vp)]),
    ("evdb_store_destroy", None, [_vp]),
    ("evdb_store_stats", _i, [_vp, C.POINTER(Stats)]),
    ("evdb_store_set_plan", _i, [_vp, _i]),
    ("evdb_store_flush", _i, [_vp]),
    ("evdb_store_profile", _i, [_vp, _i]),
    ("evdb_store_profile_read", _i, [_vp, _pi32, _pd]),
    ("evdb_store_upsert_f64", _i, [_vp, _u32, _pd, _i]),
    ("evdb_store_upsert_f32", _i, [_vp, _u32, _pf, _i]),
    ("evdb_store_append_f64", _i, [_vp, _pd, _u64, _i, C.POINTER(C.c_uint64)]),
    ("evdb_store_append_f32", _i, [_vp, _pf, _u64, _i, C.POINTER(C.c_uint64)]),
    ("evdb_store_bulk_load_f32", _i, [_vp, _pf, _u64, _i]),
    ("evdb_store_bulk_load_f64", _i, [_vp, _pd, _u64, _i]),
    ("evdb_store_bulk_load_codes", _i, [_vp, _pu8, _pd, _pd, _u64, _i]),
    ("evdb_store_delete", _i, [_vp, _u32, _pi64]),
    ("evdb_store_get_f64", _i, [_vp, _u32, _pd, _i]),
    ("evdb_store_get_codes", _i, [_vp, _u32, _pu8, _pd, _pd]),
    ("evdb_store_fill_synthetic", _i, [_vp, _u64, _u64, _u64, _i]),
    ("evdb_store_search_f64", _i, [_vp, _pd, _i, _i, _i, _i, _pu32, _pd, _pi32]),
    ("evdb_store_search_f32", _i, [_vp, _pf, _i, _i, _i, _i, _pu32, _pd, _pi32]),
    ("evdb_store_search_dev", _i, [_vp, _vp, _i, _i, _i, _i, _u64, _vp, _vp, _vp, _vp, _vp]),
    ("evdb_store_search_dev_ex", _i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(SearchOpts), _vp, _vp, _vp, _vp, _vp]),
    ("evdb_merge_topk_dev", _i, [_i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    ("evdb_merge_topk_packed_dev", _i, [_i, _vp, _i, _i, _i, _vp, _vp]),
    ("evdb_exchange_create", _i, [_i, _i, _i, _u64, C.POINTER(_vp), _vp]),
    ("evdb_exchange_connect", _i, [_vp, _vp]),
    ("evdb_exchange_connect_ptrs", _i, [_vp, C.POINTER(_vp)]),
    ("evdb_exchange_mailbox", _vp, [_vp]),
    ("evdb_exchange_push", _i, [_vp, _vp, _i, _i, _vp]),
    ("evdb_exchange_merge", _i, [_vp, _i, _i, _vp, _vp]),
    ("evdb_exchange_destroy", None, [_vp]),
    ("evdb_store_search_sharded_phase1", _i, [_vp, _vp, _vp, _i, _i, _i, _i, _u64, _u64, _vp]),
    ("evdb_store_search_sharded_phase2", _i, [_vp, _vp, _vp, _vp, _i, _i, _i, _u64, _vp]),
    ("evdb_store_search_sharded_phase3", _i, [_vp, _vp, _i, _i, _i, _u64, _vp, _vp]),
    ("evdb_quantize_8bit", _i, [_i, _pd, _u64, _i, _pu8, _pd, _pd, _pd, _pu8]),
    ("evdb_quantize_4bit", _i, [_i, _pd, _u64, _i, _pu8, _pd, _pd, _pd, _pu8]),
    ("evdb_dequantize_8bit", _i, [_i, _pu8, _pd, _pd, _u64, _i, _pd]),
    ("evdb_dequantize_4bit", _i, [_i, _pu8, _pd, _pd, _u64, _i, _pd]),
    ("evdb_vector_utils_f64", _i, [_i, _i, _pd, _pd, _u64, _i, _pd]),
    ("evdb_debug_scan_tile_plan", _i, [_i, _i, _i, _u64, _i, _pi32]),
    ("evdb_debug_quant_dots", _i, [_vp, _pd, _i, _pu32, _i, _pi64, _pi32, _pi32, _pi32]),
]

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library or fail loudly (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m erlvectordb_b200.build` "
                "(nvcc, sm_100a). erlvectordb_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(L, name)  # AttributeError if the ABI lost a symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc: int, where: str) -> None:
    if rc != OK:
        raise EvdbError(rc, where)
