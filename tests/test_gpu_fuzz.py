"""Seeded differential sweep over odd shapes: the tcgen05 GEMM plan, the scan plan and the strict
oracle must agree bit for bit on ids and fp64 distances (ragged row counts, dimensions that are not
multiples of anything, k around the window sizes, batches around the 128-query block and the
8-query CTA of the select kernel)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(20240917)
    out = []
    for i in range(14):
        n = int(rng.integers(257, 30_000))
        d = int(rng.choice([1, 3, 17, 63, 64, 65, 100, 129, 200, 257, 300]))
        B = int(rng.choice([16, 17, 31, 64, 127, 128, 129, 200, 257]))
        k = int(rng.choice([1, 5, 10, 26, 27, 50, 51, 100, 102]))
        metric = ["cosine", "euclidean"][i % 2]
        out.append((n, d, B, k, metric, int(rng.integers(1, 1 << 30))))
    return out


@pytest.mark.parametrize("n,d,B,k,metric,seed", _cases())
def test_gemm_scan_oracle_agree(native, oracle, n, d, B, k, metric, seed):
    from erlvectordb_b200.device_store import DeviceStore
    rng = np.random.default_rng(seed)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[rng.integers(0, n, size=3)] *= 50.0          # a few large-norm rows
    if n > 300:
        rows[7] = rows[300]                            # an exact duplicate (tie broken by slot)
    st = DeviceStore(dtype="f32", device=0)
    st.bulk_load(rows)
    qs = rng.standard_normal((B, d))
    qs[B // 2] = rows[n // 2].astype(np.float64) * 1.5  # a query colinear with a stored row
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, k, metric)
    assert st.stats()["last_plan"] == native.PLAN_GEMM
    st.set_plan("scan")
    ss, sd, sc = st.search(qs, k, metric)
    assert np.array_equal(gc, sc) and np.array_equal(gs, ss) and np.array_equal(gd, sd)
    ref = rows.astype(np.float64)
    for b in (0, B // 2, B - 1):
        r, dd = oracle.search(ref, qs[b], k, metric)
        kk = min(k, n)
        assert gc[b] == kk
        assert gs[b, :kk].tolist() == r.tolist() and gd[b, :kk].tolist() == dd.tolist(), (b,)
    st.close()
