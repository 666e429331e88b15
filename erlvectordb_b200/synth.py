"""Counter-based synthetic vectors, numpy edition for host-side
queries: value(seed, idx) = ((splitmix64(seed ^ idx) >> 40) - 2^23) * 2^-23, idx = row*d + col.
A 24-bit grid in [-1, 1): every value is exact in fp32 and fp64.  The same function is
implemented on the device (csrc/common.cuh synth_value) for corpora that never cross PCIe."""
from __future__ import annotations

import numpy as np

SEED_CORPUS = 0x5EED0001
SEED_QUERY = 0x5EED0002


def synth(seed: int, row0: int, nrows: int, d: int, dtype=np.float64) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = (np.arange(row0 * d, (row0 + nrows) * d, dtype=np.uint64)).reshape(nrows, d)
        z = np.uint64(seed) ^ idx
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    m = (z >> np.uint64(40)).astype(np.int64) - 8388608
    return (m.astype(np.float64) * (1.0 / 8388608.0)).astype(dtype)
