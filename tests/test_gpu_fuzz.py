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


def _small_cases():
    rng = np.random.default_rng(77)
    out = []
    for i in range(18):
        n = int(rng.choice([1, 2, 9, 31, 100, 257, 1000, 4097, 20_000]))
        d = int(rng.choice([1, 2, 5, 16, 33, 96, 128, 255, 256, 257, 400]))
        B = int(rng.choice([1, 1, 2, 3, 5, 8, 9, 13]))
        k = int(rng.choice([1, 3, 10, 40, 100]))
        out.append((n, d, B, k, ["cosine", "euclidean", "manhattan"][i % 3], ["f32", "bf16"][(i // 3) % 2], int(rng.integers(1, 1 << 30))))
    return out


@pytest.mark.parametrize("n,d,B,k,metric,dtype,seed", _small_cases())
def test_small_store_and_multi_query_paths_agree_with_the_exhaustive_plan(native, oracle, n, d, B, k, metric, dtype, seed):
    """The one-launch small-store search (B <= 2, d <= 256) and the multi-query float scan (B >= 2) over odd
    shapes -- fewer rows than k, one-element vectors, ragged query groups, duplicates, a zero row -- against
    the exhaustive fp64 plan of the same store, and for fp32 stores against the strict oracle."""
    from erlvectordb_b200.device_store import DeviceStore
    rng = np.random.default_rng(seed)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    if n > 3:
        rows[1] = rows[3]            # exact duplicate: tie broken by slot
        rows[2] = 0.0                # zero norm: cosine distance 1.0
    qs = rng.standard_normal((B, d))
    st = DeviceStore(dtype=dtype, device=0)
    try:
        st.bulk_load(rows)
        got = st.search(qs, k, metric)
        assert st.stats()["last_plan"] in (native.PLAN_SCAN, native.PLAN_GEMM, native.PLAN_EXACT)
        st.set_plan("exact")
        want = st.search(qs, k, metric)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        if dtype == "f32":
            ref = rows.astype(np.float64)
            kk = min(k, n)
            for b in range(B):
                r, dd = oracle.search(ref, qs[b], k, metric)
                assert got[2][b] == kk and got[0][b, :kk].tolist() == r.tolist() and got[1][b, :kk].tolist() == dd.tolist()
    finally:
        st.close()
