"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see evdb_oracle.c).

Three tiers, all restating the reference Erlang text they cite:

* ``strict``  -- ctypes binding of ``libevdb_oracle.so`` (scalar fp64 C, the
  exact operation order of src/vector_store.erl:227-252 etc.).
* ``py_*``    -- pure-Python twins of the same functions, used to pin the C
  against the known-answer vectors independently (small inputs only).
* ``bulk_*``  -- numpy fp64 chunked evaluation for parity sweeps with many
  queries; agrees with the strict tier to ~1e-13 (different summation order),
  which is far inside the 1e-5 parity tolerance.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libevdb_oracle.so")

COSINE, EUCLIDEAN, MANHATTAN = 0, 1, 2
METRICS = {"cosine": COSINE, "euclidean": EUCLIDEAN, "manhattan": MANHATTAN}

SEED_CORPUS = 0x5EED0001
SEED_QUERY = 0x5EED0002


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "evdb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libevdb_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.evo_dot.restype = C.c_double
        L.evo_dot.argtypes = [dp, dp, C.c_int]
        L.evo_norm.restype = C.c_double
        L.evo_norm.argtypes = [dp, C.c_int]
        for name in ("evo_cosine_distance", "evo_cosine_similarity",
                     "evo_euclidean_distance", "evo_manhattan_distance"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [dp, dp, C.c_int]
        L.evo_search.restype = C.c_int64
        L.evo_search.argtypes = [dp, C.c_int64, C.c_int, dp, C.c_int64, C.c_int,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_int64), dp]
        L.evo_search_f32.restype = C.c_int64
        L.evo_search_f32.argtypes = [C.POINTER(C.c_float), C.c_int64, C.c_int, dp,
                                     C.c_int64, C.c_int, C.POINTER(C.c_int64), dp]
        L.evo_distances.restype = None
        L.evo_distances.argtypes = [dp, C.c_int64, C.c_int, dp, C.c_int, dp]
        L.evo_quantize_8bit.restype = C.c_int
        L.evo_quantize_8bit.argtypes = [dp, C.c_int, C.POINTER(C.c_uint8), dp, dp, dp]
        L.evo_quantize_4bit.restype = C.c_int
        L.evo_quantize_4bit.argtypes = [dp, C.c_int, C.POINTER(C.c_uint8), dp, dp, dp]
        L.evo_dequantize_8bit.restype = None
        L.evo_dequantize_8bit.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_double,
                                          C.c_double, dp]
        L.evo_dequantize_4bit.restype = None
        L.evo_dequantize_4bit.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_double,
                                          C.c_double, dp]
        L.evo_synth_value.restype = C.c_double
        L.evo_synth_value.argtypes = [C.c_uint64, C.c_uint64]
        L.evo_synth_fill_f64.restype = None
        L.evo_synth_fill_f64.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int, dp]
        L.evo_synth_fill_f32.restype = None
        L.evo_synth_fill_f32.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int,
                                         C.POINTER(C.c_float)]
        L.evo_synth_fill_f32_mt.restype = None
        L.evo_synth_fill_f32_mt.argtypes = [C.c_uint64, C.c_int64, C.c_int,
                                            C.POINTER(C.c_float), C.c_int]
        L.evo_search_f32_replicas.restype = C.c_int
        L.evo_search_f32_replicas.argtypes = [C.POINTER(C.c_float), C.c_int64, C.c_int,
                                              dp, C.c_int, C.c_int64, C.c_int,
                                              C.POINTER(C.c_int64), dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64))


# ----------------------------------------------------------------- strict tier
def distance(q, v, metric="cosine") -> float:
    q, v = _f64(q), _f64(v)
    f = {COSINE: lib().evo_cosine_distance, EUCLIDEAN: lib().evo_euclidean_distance,
         MANHATTAN: lib().evo_manhattan_distance}[METRICS.get(metric, metric)]
    return f(_dp(q), _dp(v), len(q))


def distances(rows, q, metric="cosine") -> np.ndarray:
    rows, q = _f64(rows), _f64(q)
    n, d = rows.shape
    out = np.empty(n, dtype=np.float64)
    lib().evo_distances(_dp(rows), n, d, _dp(q), METRICS.get(metric, metric), _dp(out))
    return out


def search(rows, q, k, metric="cosine", ranks=None):
    """perform_search/3 (src/vector_store.erl:227-236): (row indices, distances)."""
    rows, q = _f64(rows), _f64(q)
    n, d = rows.shape if rows.ndim == 2 else (0, len(q))
    kk = max(int(k), 0)
    out_rows = np.empty(max(min(kk, n), 1), dtype=np.int64)
    out_dist = np.empty(max(min(kk, n), 1), dtype=np.float64)
    rk = None
    if ranks is not None:
        rk_arr = np.ascontiguousarray(np.asarray(ranks, dtype=np.uint64))
        rk = rk_arr.ctypes.data_as(C.POINTER(C.c_uint64))
    m = lib().evo_search(_dp(rows), n, d, _dp(q), int(k), METRICS.get(metric, metric), rk,
                         out_rows.ctypes.data_as(C.POINTER(C.c_int64)), _dp(out_dist))
    if m < 0:
        raise ValueError("function_clause: lists:sublist/2 with negative K")
    return out_rows[:m].copy(), out_dist[:m].copy()


def id_ranks(ids):
    """Erlang term order for binary ids == Python bytes order (bytewise, then length)."""
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    ranks = [0] * len(ids)
    for r, i in enumerate(order):
        ranks[i] = r
    return ranks


def quantize_8bit(v):
    v = _f64(v)
    codes = np.empty(len(v), dtype=np.uint8)
    mn, mx, sc = C.c_double(), C.c_double(), C.c_double()
    rc = lib().evo_quantize_8bit(_dp(v), len(v), codes.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 C.byref(mn), C.byref(mx), C.byref(sc))
    if rc != 0:
        raise ArithmeticError("badarith")  # vector_compression.erl:62-64
    return codes, mn.value, mx.value, sc.value


def quantize_4bit(v):
    v = _f64(v)
    packed = np.empty((len(v) + 1) // 2, dtype=np.uint8)
    mn, mx, sc = C.c_double(), C.c_double(), C.c_double()
    rc = lib().evo_quantize_4bit(_dp(v), len(v), packed.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 C.byref(mn), C.byref(mx), C.byref(sc))
    if rc != 0:
        raise ArithmeticError("badarith")
    return packed, mn.value, mx.value, sc.value


def dequantize_8bit(codes, mn, scale):
    codes = np.ascontiguousarray(np.asarray(codes, dtype=np.uint8))
    out = np.empty(len(codes), dtype=np.float64)
    lib().evo_dequantize_8bit(codes.ctypes.data_as(C.POINTER(C.c_uint8)), len(codes),
                              mn, scale, _dp(out))
    return out


def dequantize_4bit(packed, d, mn, scale):
    packed = np.ascontiguousarray(np.asarray(packed, dtype=np.uint8))
    out = np.empty(d, dtype=np.float64)
    lib().evo_dequantize_4bit(packed.ctypes.data_as(C.POINTER(C.c_uint8)), d, mn, scale,
                              _dp(out))
    return out


def synth_f64(seed, row0, nrows, d) -> np.ndarray:
    out = np.empty((nrows, d), dtype=np.float64)
    lib().evo_synth_fill_f64(seed, row0, nrows, d, _dp(out))
    return out


def synth_f32(seed, row0, nrows, d, threads=1) -> np.ndarray:
    out = np.empty((nrows, d), dtype=np.float32)
    p = out.ctypes.data_as(C.POINTER(C.c_float))
    if threads > 1 and row0 == 0:
        lib().evo_synth_fill_f32_mt(seed, nrows, d, p, threads)
    else:
        lib().evo_synth_fill_f32(seed, row0, nrows, d, p)
    return out


def search_f32_replicas(rows_f32, queries, k, metric="cosine"):
    """T = len(queries) threads, each a full reference search (CPU baseline)."""
    rows_f32 = np.ascontiguousarray(rows_f32, dtype=np.float32)
    queries = _f64(queries)
    n, d = rows_f32.shape
    nq = queries.shape[0]
    out_rows = np.zeros((nq, k), dtype=np.int64)
    out_dist = np.zeros((nq, k), dtype=np.float64)
    rc = lib().evo_search_f32_replicas(rows_f32.ctypes.data_as(C.POINTER(C.c_float)), n, d,
                                       _dp(queries), nq, k, METRICS.get(metric, metric),
                                       out_rows.ctypes.data_as(C.POINTER(C.c_int64)),
                                       _dp(out_dist))
    if rc != 0:
        raise RuntimeError("oracle search failed")
    return out_rows, out_dist


# ------------------------------------------------------------ pure-Python twin
def py_dot(a, b):
    s = 0
    for x, y in zip(a, b):
        s = s + x * y
    return s


def py_norm(v):
    s = 0
    for x in v:
        s = s + x * x
    return math.sqrt(s)


def py_cosine_distance(q, v):
    dot, n1, n2 = py_dot(q, v), py_norm(q), py_norm(v)
    if n1 == 0.0 or n2 == 0.0:
        return 1.0
    return 1.0 - (dot / (n1 * n2))


def py_euclidean(a, b):
    return py_norm([x - y for x, y in zip(a, b)])


def py_manhattan(a, b):
    s = 0
    for x, y in zip(a, b):
        s = s + abs(x - y)
    return s


def py_round_half_away(t):
    return int(math.floor(abs(t) + 0.5)) * (1 if t >= 0 else -1)


def py_quantize(v, levels):
    mn, mx = v[0], v[0]
    for x in v[1:]:
        mn, mx = min(x, mn), max(x, mx)
    scale = (mx - mn) / float(levels)
    if scale == 0.0:
        raise ArithmeticError("badarith")
    return [py_round_half_away((x - mn) / scale) for x in v], mn, mx, scale


def py_pack_4bit(codes):
    out = bytearray()
    for i in range(0, len(codes), 2):
        lo = codes[i + 1] if i + 1 < len(codes) else 0
        out.append((codes[i] << 4) | lo)
    return bytes(out)


def py_search(entries, q, k, metric=py_cosine_distance):
    """entries: list of (id_bytes, vector).  Mirrors perform_search/3."""
    if k < 0:
        raise ValueError("function_clause")
    dist = sorted((metric(q, v), i) for i, v in entries)
    return [(i, dd) for dd, i in dist[:k]]


# ------------------------------------------------------------------- bulk tier
def bulk_distances(rows, queries, metric="cosine") -> np.ndarray:
    """fp64 numpy distances, shape (B, N).  Same formulas, numpy summation order."""
    rows = np.asarray(rows, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    m = METRICS.get(metric, metric)
    if m == COSINE:
        dots = queries @ rows.T
        nq = np.sqrt((queries * queries).sum(axis=1))[:, None]
        nv = np.sqrt((rows * rows).sum(axis=1))[None, :]
        den = nq * nv
        with np.errstate(divide="ignore", invalid="ignore"):
            out = 1.0 - dots / den
        out[np.broadcast_to(den == 0.0, out.shape)] = 1.0
        return out
    out = np.empty((queries.shape[0], rows.shape[0]), dtype=np.float64)
    for b in range(queries.shape[0]):
        diff = rows - queries[b][None, :]
        out[b] = np.sqrt((diff * diff).sum(axis=1)) if m == EUCLIDEAN else np.abs(diff).sum(axis=1)
    return out


def bulk_search(rows, queries, k, metric="cosine", row0=0):
    """Top-k per query by (distance, row).  Returns (B,k) rows and distances."""
    dist = bulk_distances(rows, queries, metric)
    n = dist.shape[1]
    kk = min(k, n)
    idx = np.lexsort((np.broadcast_to(np.arange(n), dist.shape), dist), axis=1)[:, :kk]
    return idx + row0, np.take_along_axis(dist, idx, axis=1)


def bulk_search_synth(seed, n, d, queries, k, metric="cosine", chunk=65536):
    """Top-k over a synthetic corpus generated chunk by chunk (no N x d array)."""
    queries = np.asarray(queries, dtype=np.float64)
    B = queries.shape[0]
    best_d = np.full((B, 0), np.inf)
    best_i = np.zeros((B, 0), dtype=np.int64)
    for r0 in range(0, n, chunk):
        cnt = min(chunk, n - r0)
        rows = synth_f64(seed, r0, cnt, d)
        i, dd = bulk_search(rows, queries, k, metric, row0=r0)
        best_d = np.concatenate([best_d, dd], axis=1)
        best_i = np.concatenate([best_i, i], axis=1)
        order = np.lexsort((best_i, best_d), axis=1)[:, :k]
        best_d = np.take_along_axis(best_d, order, axis=1)
        best_i = np.take_along_axis(best_i, order, axis=1)
    return best_i, best_d
