"""Mirror of the reference's ``vector_compression`` for the two scan formats
(reference src/vector_compression.erl:44-91,166-204), computed on the device in
fp64 so codes are bit-identical to the Erlang arithmetic.

Only quantization_8bit / quantization_4bit are in scope;
the reference's placeholder PCA / zlib / LZ4 / PQ codecs are storage-only and are
answered with ``{error, {unsupported_algorithm, A}}`` here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

SUPPORTED = ("quantization_8bit", "quantization_4bit")


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def compress_batch(vectors, algorithm, device=0):
    """{ok, [CompressedVector]} | {error, batch_compression_failed} (:96-109)."""
    if algorithm not in SUPPORTED:
        return ("error", ("unsupported_algorithm", algorithm))
    rows = np.ascontiguousarray(vectors, dtype=np.float64)
    if rows.ndim != 2 or rows.shape[1] == 0:
        return ("error", "batch_compression_failed")
    n, d = rows.shape
    eight = algorithm == "quantization_8bit"
    nb = d if eight else (d + 1) // 2
    codes = np.empty((n, nb), dtype=np.uint8)
    mins, maxs, scales = (np.empty(n, dtype=np.float64) for _ in range(3))
    ok = np.empty(n, dtype=np.uint8)
    fn = N.lib().evdb_quantize_8bit if eight else N.lib().evdb_quantize_4bit
    N.check(fn(device, _p(rows, C.c_double), n, d, _p(codes, C.c_uint8), _p(mins, C.c_double),
               _p(maxs, C.c_double), _p(scales, C.c_double), _p(ok, C.c_uint8)), "evdb_quantize")
    if not ok.all():
        return ("error", "batch_compression_failed")
    out = []
    for i in range(n):
        meta = {"min": float(mins[i]), "max": float(maxs[i]), "scale": float(scales[i])}
        if not eight:
            meta["length"] = d
        out.append({"algorithm": algorithm, "data": codes[i].tobytes(), "metadata": meta})
    return ("ok", out)


def compress_vector(vector, algorithm, device=0):
    """{ok, #{algorithm, data, metadata}} | {error, _} (:44-65)."""
    if algorithm not in SUPPORTED:
        return ("error", ("unsupported_algorithm", algorithm))
    if len(vector) == 0:
        return ("error", ("compression_failed", "error", "badarg"))  # hd([]) in find_min_max
    r = compress_batch([vector], algorithm, device)
    if r[0] == "error":
        # Max == Min divides by 0.0 in the reference (:169,:188)
        return ("error", ("compression_failed", "error", "badarith"))
    return ("ok", r[1][0])


def decompress_batch(compressed, options=None, device=0):
    if not compressed:
        return ("ok", [])
    alg = compressed[0]["algorithm"]
    if alg not in SUPPORTED or any(c["algorithm"] != alg for c in compressed):
        return ("error", "batch_decompression_failed")
    eight = alg == "quantization_8bit"
    n = len(compressed)
    d = len(compressed[0]["data"]) if eight else compressed[0]["metadata"]["length"]
    codes = np.frombuffer(b"".join(c["data"] for c in compressed), dtype=np.uint8).copy()
    mins = np.array([c["metadata"]["min"] for c in compressed], dtype=np.float64)
    scales = np.array([c["metadata"]["scale"] for c in compressed], dtype=np.float64)
    out = np.empty((n, d), dtype=np.float64)
    fn = N.lib().evdb_dequantize_8bit if eight else N.lib().evdb_dequantize_4bit
    N.check(fn(device, _p(codes, C.c_uint8), _p(mins, C.c_double), _p(scales, C.c_double), n, d,
               _p(out, C.c_double)), "evdb_dequantize")
    return ("ok", [row.tolist() for row in out])


def decompress_vector(compressed, options=None, device=0):
    if compressed.get("algorithm") not in SUPPORTED:
        return ("error", ("unsupported_algorithm", compressed.get("algorithm")))
    r = decompress_batch([compressed], options, device)
    return ("ok", r[1][0]) if r[0] == "ok" else r


def get_compression_ratio(original_vector, compressed):
    """OriginalSize (4 bytes per float) / byte_size(data)  (:122-126)."""
    return (len(original_vector) * 4) / len(compressed["data"])


def get_supported_algorithms():
    return list(SUPPORTED)
