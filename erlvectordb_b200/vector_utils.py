"""Mirror of the reference's vector_utils module (src/vector_utils.erl:28-57): the pairwise
functions, same names and results, computed by libevdb_b200 on the device in the reference's
operation order (bit-equal to the Erlang fp64 result).  cosine_similarity/2 (:28-36) is the
similarity -- search uses the DISTANCE form of vector_store:cosine_distance/2.

Each function takes two vectors (or two (n, d) arrays: n pairs) and returns a float (or n floats).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

COSINE_SIMILARITY, COSINE_DISTANCE, EUCLIDEAN, MANHATTAN, DOT, NORM = range(6)


def _pairs(op: int, a, b=None, device: int = 0):
    x = np.ascontiguousarray(a, dtype=np.float64)
    single = x.ndim == 1
    x = x.reshape(1, -1) if single else x
    y = None
    if b is not None:
        y = np.ascontiguousarray(b, dtype=np.float64).reshape(x.shape[0], -1)
        if y.shape != x.shape:
            raise ValueError("vectors must have the same length")   # the reference's lists:zip/2 raises too
    out = np.empty(x.shape[0], dtype=np.float64)
    dp = C.POINTER(C.c_double)
    N.check(N.lib().evdb_vector_utils_f64(device, op, x.ctypes.data_as(dp), y.ctypes.data_as(dp) if y is not None else None,
                                          x.shape[0], x.shape[1], out.ctypes.data_as(dp)), "evdb_vector_utils_f64")
    return float(out[0]) if single else out


def cosine_similarity(v1, v2, device=0):
    return _pairs(COSINE_SIMILARITY, v1, v2, device)


def euclidean_distance(v1, v2, device=0):
    return _pairs(EUCLIDEAN, v1, v2, device)


def manhattan_distance(v1, v2, device=0):
    return _pairs(MANHATTAN, v1, v2, device)


def dot_product(v1, v2, device=0):
    return _pairs(DOT, v1, v2, device)


def vector_norm(v, device=0):
    return _pairs(NORM, v, None, device)
