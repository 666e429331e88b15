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
    st.search(qs, 10, "euclidean")          # no GEMM form implemented for L2 yet: scan
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
