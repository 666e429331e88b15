"""The tcgen05 GEMM plan (query batches) must return exactly what the scan plan and the
oracle return: the GEMM only proposes candidates, select.cu re-ranks them in fp64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(native, n, d, seed=None, rows=None):
    from erlvectordb_b200.device_store import DeviceStore
    st = DeviceStore(dtype="f32", gemm_shadow=True)
    if rows is None:
        st.fill_synthetic(seed, n, d)
    else:
        st.bulk_load(rows)
    return st


@pytest.mark.parametrize("n,d,B,k", [
    (50_000, 768, 128, 10),     # exact tile multiples in B
    (100_003, 768, 200, 10),    # ragged rows (last tile partial) and ragged batch
    (30_000, 100, 40, 10),      # K not a multiple of 64 (TMA zero-fills the tail), one query block
    (70_001, 128, 700, 5),      # several query blocks and sweeps
])
def test_gemm_plan_equals_scan_plan_and_oracle(native, oracle, n, d, B, k):
    st = _mk(native, n, d, seed=oracle.SEED_CORPUS)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, k, "cosine")
    assert st.stats()["last_plan"] == native.PLAN_GEMM
    esc = st.stats()["escalations"]
    st.set_plan("scan")
    ss, sd, sc = st.search(qs, k, "cosine")
    assert st.stats()["last_plan"] == native.PLAN_SCAN
    assert np.array_equal(gc, sc) and np.array_equal(gs, ss)
    assert np.array_equal(gd, sd)  # bit-identical fp64 distances from either plan
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d)
    for b in (0, B // 2, B - 1):
        r, dd = oracle.search(rows, qs[b], k, "cosine")
        assert gs[b].tolist() == r.tolist() and gd[b].tolist() == dd.tolist()
    assert esc <= max(2, B // 50), f"{esc} of {B} candidate windows had to be escalated"
    st.close()


def test_auto_plan_picks_gemm_for_batches_and_scan_for_single(native, oracle):
    st = _mk(native, 20_000, 256, seed=oracle.SEED_CORPUS)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 64, 256)
    a = st.search(qs, 10, "cosine")
    assert st.stats()["last_plan"] == native.PLAN_GEMM
    b = st.search(qs[:1], 10, "cosine")
    assert st.stats()["last_plan"] == native.PLAN_SCAN
    assert a[0][0].tolist() == b[0][0].tolist() and a[1][0].tolist() == b[1][0].tolist()
    st.search(qs, 10, "euclidean")          # L2 has a GEMM form (row norms ride in the K dimension)
    assert st.stats()["last_plan"] == native.PLAN_GEMM
    st.search(qs, 10, "manhattan")          # L1 has none: scan
    assert st.stats()["last_plan"] == native.PLAN_SCAN
    st.close()


def test_gemm_with_scaled_and_degenerate_rows(native, oracle):
    """Rows of wildly different norms, a zero row and duplicates: the fp16 operands are unit
    vectors, so scale cannot overflow them; ties are still broken by slot."""
    rng = np.random.default_rng(4)
    n, d = 4096, 192
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[::7] *= 1e4
    rows[1::7] *= 1e-4
    rows[100] = 0.0
    rows[200] = rows[300]
    st = _mk(native, n, d, rows=rows)
    qs = rng.standard_normal((96, d))
    qs[5] = rows[300].astype(np.float64) * 3.0
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, 10, "cosine")
    ref = rows.astype(np.float64)
    for b in (0, 5, 50, 95):
        r, dd = oracle.search(ref, qs[b], 10, "cosine")
        assert gs[b].tolist() == r.tolist() and gd[b].tolist() == dd.tolist()
    assert gs[5, 0] == 200 and gs[5, 1] == 300
    st.close()


def test_upsert_keeps_shadow_current(native, oracle):
    rng = np.random.default_rng(6)
    n, d = 3000, 64
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = _mk(native, n, d, rows=rows)
    q = rng.standard_normal((32, d))
    target = (q[3] * 2.0).astype(np.float32)
    assert st.upsert(17, target.astype(np.float64)) == 0   # overwrite
    assert st.upsert(n, target.astype(np.float64)) == 0    # append
    rows = np.vstack([rows, target[None, :]])
    rows[17] = target
    st.set_plan("gemm")
    gs, gd, gc = st.search(q, 4, "cosine")
    r, dd = oracle.search(rows.astype(np.float64), q[3], 4, "cosine")
    assert gs[3].tolist() == r.tolist() and gd[3].tolist() == dd.tolist()
    assert set(gs[3, :2].tolist()) == {17, n}
    st.close()


@pytest.mark.parametrize("n,d,B,k", [
    (60_000, 128, 256, 100),    # BASELINE configs[2] shape (10M x 128, k = 100) scaled down: 128-key windows
    (100_003, 96, 130, 10),     # ragged rows / batch, K + 3 norm columns crossing a 64-column block
    (40_000, 61, 64, 25),       # odd dimension: K = 64 exactly after the norm columns
])
def test_euclidean_gemm_equals_scan_plan_and_oracle(native, oracle, n, d, B, k):
    """||q - v||^2 = ||q||^2 - 2(q.v - ||v||^2/2): the row norms are three extra fp16 columns of the
    GEMM operand, candidates are re-ranked in exact fp64 (sqrt of the left-to-right sum)."""
    st = _mk(native, n, d, seed=oracle.SEED_CORPUS)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, k, "euclidean")
    assert st.stats()["last_plan"] == native.PLAN_GEMM
    esc = st.stats()["escalations"]
    st.set_plan("scan")
    ss, sd, sc = st.search(qs, k, "euclidean")
    assert np.array_equal(gc, sc) and np.array_equal(gs, ss)
    assert np.array_equal(gd, sd)
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d)
    for b in (0, B // 2, B - 1):
        r, dd = oracle.search(rows, qs[b], k, "euclidean")
        assert gs[b].tolist() == r.tolist() and gd[b].tolist() == dd.tolist()
    assert esc <= max(2, B // 50), f"{esc} of {B} candidate windows had to be escalated"
    st.close()


def test_euclidean_gemm_scales_and_near_duplicates(native, oracle):
    """Row norms spanning 1e-3..1e3, a zero row, queries that ARE stored rows (distance 0: the
    fp16 window cannot be proven there, the query must escalate and still come back exact),
    and an upsert with a larger norm than the operand scale was chosen for."""
    rng = np.random.default_rng(11)
    n, d = 8192, 160
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[::5] *= 1e2
    rows[1::5] *= 1e-3
    rows[77] = 0.0
    st = _mk(native, n, d, rows=rows)
    qs = rng.standard_normal((64, d))
    qs[3] = rows[1000].astype(np.float64)
    qs[4] = rows[5].astype(np.float64) * 1.0001
    qs[9] = 0.0
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, 8, "euclidean")
    ref = rows.astype(np.float64)
    for b in (0, 3, 4, 9, 63):
        r, dd = oracle.search(ref, qs[b], 8, "euclidean")
        assert gs[b].tolist() == r.tolist() and gd[b].tolist() == dd.tolist(), b
    assert gs[3, 0] == 1000 and gd[3, 0] == 0.0
    big = (rng.standard_normal(d) * 1e4).astype(np.float32)
    assert st.upsert(n, big.astype(np.float64)) == 0        # outgrows sigma: column is rebuilt
    assert st.upsert(12, (big * 0.5).astype(np.float64)) == 0
    rows = np.vstack([rows, big[None, :]])
    rows[12] = big * 0.5
    q2 = np.stack([big.astype(np.float64) * 0.9, qs[0]])
    q2 = np.vstack([q2, rng.standard_normal((30, d))])
    gs, gd, gc = st.search(q2, 4, "euclidean")
    assert st.stats()["last_plan"] == native.PLAN_GEMM
    for b in (0, 1, 17):
        r, dd = oracle.search(rows.astype(np.float64), q2[b], 4, "euclidean")
        assert gs[b].tolist() == r.tolist() and gd[b].tolist() == dd.tolist(), b
    assert gs[0, 0] == n
    st.delete(3)                                              # swap-with-last keeps the column coherent
    rows[3] = rows[-1]
    rows = rows[:-1]
    gs, gd, gc = st.search(q2, 4, "euclidean")
    for b in (0, 5):
        r, dd = oracle.search(rows.astype(np.float64), q2[b], 4, "euclidean")
        assert gs[b].tolist() == r.tolist() and gd[b].tolist() == dd.tolist(), b
    st.close()


def test_gemm_large_k_window(native, oracle):
    """k = 100 on the cosine GEMM plan (128-key candidate windows, 4 keys per lane in the flush sort)."""
    n, d, B, k = 50_000, 256, 96, 100
    st = _mk(native, n, d, seed=oracle.SEED_CORPUS)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, k, "cosine")
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d)
    for b in (0, 50, B - 1):
        r, dd = oracle.search(rows, qs[b], k, "cosine")
        assert gs[b].tolist() == r.tolist() and gd[b].tolist() == dd.tolist()
    st.close()


def test_large_batch_with_a_non_finite_query_is_rejected(native, oracle):
    """Batches are validated on the host while the device already works on them; a NaN/Inf anywhere
    must still come back as invalid_vector_format, and the store must stay usable."""
    n, d, B = 30_000, 256, 300
    st = _mk(native, n, d, seed=oracle.SEED_CORPUS)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    bad = qs.copy()
    bad[211, 17] = np.nan
    assert st.search(bad, 10, "cosine") == native.E_BAD_VECTOR
    bad[211, 17] = np.inf
    assert st.search(bad, 10, "euclidean") == native.E_BAD_VECTOR
    gs, gd, gc = st.search(qs, 10, "cosine")
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d)
    r, dd = oracle.search(rows, qs[211], 10, "cosine")
    assert gs[211].tolist() == r.tolist() and gd[211].tolist() == dd.tolist()
    st.close()


def test_batch_larger_than_one_launch_is_sliced(native, oracle):
    """More than 8192 queries do not fit one launch's candidate buffers: search_core slices the
    batch; every slice must land in the right rows of the result."""
    n, d, B, k = 3000, 64, 8192 + 300, 5
    st = _mk(native, n, d, seed=oracle.SEED_CORPUS)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, k, "cosine")
    assert st.stats()["last_plan"] == native.PLAN_GEMM
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d)
    for b in (0, 8191, 8192, B - 1):
        r, dd = oracle.search(rows, qs[b], k, "cosine")
        assert gs[b].tolist() == r.tolist() and gd[b].tolist() == dd.tolist(), b
    assert (gc == k).all()
    st.close()


@pytest.mark.parametrize("metric,k", [("cosine", 10), ("euclidean", 100)])
def test_auto_plan_uses_tcgen05_for_a_lone_query_on_a_large_store(native, oracle, metric, k):
    """B * N * row_bytes >= 256 MB: AUTO answers even a single query through the tcgen05 candidate
    pass (2-byte operand column read once) -- the result must not depend on that choice."""
    n, d = 600_000, 128          # 307 MB of fp32 rows
    st = _mk(native, n, d, seed=oracle.SEED_CORPUS)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 3, d)
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d)
    for b in range(3):
        gs, gd, gc = st.search(qs[b:b + 1], k, metric)
        assert st.stats()["last_plan"] == native.PLAN_GEMM
        r, dd = oracle.search(rows, qs[b], k, metric)
        assert gs[0].tolist() == r.tolist() and gd[0].tolist() == dd.tolist()
    st.set_plan("scan")
    ss, sd, sc = st.search(qs, k, metric)
    st.set_plan("auto")
    gs, gd, gc = st.search(qs, k, metric)          # three queries: 921 MB, tcgen05 plan
    assert st.stats()["last_plan"] == native.PLAN_GEMM
    assert np.array_equal(gs, ss) and np.array_equal(gd, sd) and np.array_equal(gc, sc)
    st.search(qs[:1], k, "manhattan")              # no GEMM form
    assert st.stats()["last_plan"] == native.PLAN_SCAN
    st.close()
