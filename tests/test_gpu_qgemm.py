"""Query batches against quantization_8bit stores on the tensor cores (csrc/gemm_i8.cu:
tcgen05.mma.kind::i8 over the stored codes x the query's digit planes).

The plan only generates candidates; what the caller sees is the exact fp64 re-rank from the codes
(reference: vector_persistence:decompress_if_needed src/vector_persistence.erl:276-284 feeding
cosine_distance/2 src/vector_store.erl:238-246).  So every result here must be bit-equal (==) to the
exhaustive fp64 plan on the same store, to the dp4a scan plan, and -- on the winners -- to the CPU
oracle on the dequantised rows.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _store(native, dtype="u8", **kw):
    from erlvectordb_b200.device_store import DeviceStore
    return DeviceStore(dtype=dtype, **kw)


def _three_plans(native, st, qs, k):
    out = {}
    for plan, want in (("gemm", native.PLAN_GEMM), ("scan", native.PLAN_SCAN), ("exact", native.PLAN_EXACT)):
        st.set_plan(plan)
        out[plan] = st.search(qs, k, "cosine")
        assert not isinstance(out[plan], int), (plan, out[plan])
        assert st.stats()["last_plan"] == want, plan
    st.set_plan("auto")
    return out


def _all_equal(out):
    for plan in ("gemm", "scan"):
        for x, y in zip(out[plan], out["exact"]):
            assert np.array_equal(x, y), plan


@pytest.mark.parametrize("n,d,B,k", [
    (128, 16, 4, 5),            # one tile, one K-step
    (1000, 96, 64, 10),         # BASELINE configs[3] row shape: one K block of three K-steps
    (5003, 100, 130, 10),       # ragged everything: rows, dimension (zero-padded chunk), two query blocks
    (20_000, 128, 256, 100),    # 128-key windows; the last dimension the mantissa-trick conversion covers
    (6000, 200, 70, 10),        # two K blocks, int -> float by I2F
    (3001, 768, 1030, 10),      # six K blocks, nine query blocks = two sweeps with a padding block
    (700, 1536, 33, 20),
])
def test_i8_batches_equal_exhaustive_and_scan_plans(native, oracle, n, d, B, k):
    st = _store(native)
    try:
        st.fill_synthetic(oracle.SEED_CORPUS, n, d)
        qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
        out = _three_plans(native, st, qs, k)
        _all_equal(out)
        slots, dists, counts = out["gemm"]
        assert (counts == min(k, n)).all()
        for b in (0, B // 2, B - 1):      # winners against the oracle on the dequantised rows
            for j in (0, min(k, n) - 1):
                codes, mn, scale = st.get_codes(int(slots[b, j]))
                row = oracle.dequantize_8bit(codes, mn, scale)
                assert dists[b, j] == oracle.distance(qs[b], row, "cosine")
    finally:
        st.close()


def test_auto_picks_the_i8_plan_for_batches_only(native, oracle):
    st = _store(native)
    try:
        st.fill_synthetic(oracle.SEED_CORPUS, 4000, 96)
        q = oracle.synth_f64(oracle.SEED_QUERY, 0, 16, 96)
        st.search(q[:1], 10, "cosine")
        assert st.stats()["last_plan"] == native.PLAN_SCAN          # a lone query: the dp4a scan
        a = st.search(q, 10, "cosine")
        assert st.stats()["last_plan"] == native.PLAN_GEMM          # a batch: tensor cores
        st.set_plan("exact")
        e = st.search(q, 10, "cosine")
        for x, y in zip(a, e):
            assert np.array_equal(x, y)
        st.set_plan("auto")
        st.search(q, 10, "manhattan")
        assert st.stats()["last_plan"] == native.PLAN_EXACT          # no integer form: the exhaustive plan
    finally:
        st.close()


def test_i8_plan_on_real_rows_with_degenerate_ones(native, oracle):
    """Rows quantised through the reference codec (not the synthetic generator): skewed ranges, a
    constant-offset cluster, a zero query and a huge-magnitude query."""
    rng = np.random.default_rng(11)
    n, d, B, k = 3000, 48, 40, 10
    rows = rng.standard_normal((n, d)) * rng.uniform(0.01, 30.0, size=(n, 1)) + rng.uniform(-5, 5, size=(n, 1))
    rows[7] = np.linspace(100.0, 100.001, d)              # nearly constant: |min| sqrt(d) / ||y|| ~ 1
    rows[8] = np.r_[np.zeros(d - 1), 1e-3]
    qs = rng.standard_normal((B, d))
    qs[3] = 0.0                                          # cosine_distance/2 -> 1.0 for every row
    qs[4] *= 1e12
    qs[5] *= 1e-12
    qs[6] = rows[7] * 3.0
    st = _store(native)
    try:
        st.bulk_load(rows)
        out = _three_plans(native, st, qs, k)
        _all_equal(out)
        slots, dists, _ = out["gemm"]
        assert (dists[3] == 1.0).all() and slots[3].tolist() == list(range(k))
        for b in (0, 4, 5, 6):
            codes, mn, scale = st.get_codes(int(slots[b, 0]))
            assert dists[b, 0] == oracle.distance(qs[b], oracle.dequantize_8bit(codes, mn, scale), "cosine")
    finally:
        st.close()


def test_i8_plan_escalates_what_it_cannot_prove(native, oracle):
    """300 rows that quantise to (nearly) the same codes as the query's own image: the candidate keys tie
    inside the error bound, the window proof fails and the host path re-issues the query -- the
    caller still gets the exhaustive plan's answer."""
    n, d, B, k = 8000, 64, 16, 10
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    rng = np.random.default_rng(5)
    for i in range(300):
        rows[11 + 17 * i] = qs[2] * (1.0 + 1e-6 * rng.standard_normal(d))
    st = _store(native)
    try:
        st.bulk_load(rows)
        esc0 = st.stats()["escalations"]
        a = st.search(qs, k, "cosine")
        assert st.stats()["last_plan"] == native.PLAN_GEMM
        st.set_plan("exact")
        e = st.search(qs, k, "cosine")
        for x, y in zip(a, e):
            assert np.array_equal(x, y)
        assert st.stats()["escalations"] > esc0
    finally:
        st.close()


def test_i8_plan_follows_inserts_and_deletes(native, oracle):
    """The codes ARE the operand (no shadow column to keep current): after every change of the
    store the batch result equals the exhaustive plan's."""
    d, B, k = 32, 16, 5
    pool = oracle.synth_f64(oracle.SEED_CORPUS, 0, 600, d)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    st = _store(native)
    try:
        st.bulk_load(pool[:300])
        rng = np.random.default_rng(2)
        count = 300
        for step in range(12):
            if step % 3 == 2:
                st.delete(int(rng.integers(0, count))); count -= 1
            else:
                st.append(pool[300 + 20 * step: 320 + 20 * step]); count += 20
            st.set_plan("auto")
            a = st.search(qs, k, "cosine")
            assert st.stats()["last_plan"] == native.PLAN_GEMM and st.stats()["count"] == count
            st.set_plan("exact")
            e = st.search(qs, k, "cosine")
            for x, y in zip(a, e):
                assert np.array_equal(x, y)
    finally:
        st.close()


def test_multi_device_handle_uses_the_i8_plan_per_shard(native, oracle):
    from erlvectordb_b200.device_store import DeviceStore
    n, d, B, k = 9000, 96, 64, 10
    m, s = DeviceStore(dtype="u8", devices=[0, 0, 0]), DeviceStore(dtype="u8", device=0)
    try:
        m.fill_synthetic(oracle.SEED_CORPUS, n, d); s.fill_synthetic(oracle.SEED_CORPUS, n, d)
        qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
        got = m.search(qs, k, "cosine")
        assert m.stats()["last_plan"] == native.PLAN_GEMM
        s.set_plan("exact")
        for x, y in zip(got, s.search(qs, k, "cosine")):
            assert np.array_equal(x, y)
    finally:
        m.close(); s.close()


@pytest.mark.parametrize("seed", range(10))
def test_i8_plan_fuzz_against_exhaustive_plan(native, oracle, seed):
    """Random shapes, windows and data (duplicated rows = exact ties, integer-valued rows, skewed scales):
    the forced i8 plan through the escalating host entry equals the exhaustive fp64 plan bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(128, 30_000))
    d = int(rng.choice([4, 17, 64, 96, 100, 128, 129, 300, 520, 700]))
    B = int(rng.integers(1, 300))
    k = int(rng.choice([1, 10, 50, 100]))
    kind = seed % 3
    if kind == 0:
        rows = rng.standard_normal((n, d))
    elif kind == 1:
        rows = rng.integers(-6, 7, size=(n, d)).astype(np.float64)
        rows[rows.max(axis=1) == rows.min(axis=1), 0] += 1.0     # Max == Min is badarith in the reference codec
    else:
        rows = rng.standard_normal((n, d)) * np.exp(rng.uniform(-6, 6, size=(n, 1))) + rng.uniform(-3, 3, size=(n, 1))
    dup = rng.integers(0, n, size=n // 10)
    rows[dup] = rows[rng.integers(0, n, size=n // 10)]           # exact ties, broken by slot order
    qs = rng.standard_normal((B, d))
    qs[0] = rows[int(rng.integers(0, n))]
    st = _store(native)
    try:
        st.bulk_load(rows)
        st.set_plan("gemm")
        a = st.search(qs, k, "cosine")
        assert st.stats()["last_plan"] == native.PLAN_GEMM
        st.set_plan("exact")
        e = st.search(qs, k, "cosine")
        for x, y in zip(a, e):
            assert np.array_equal(x, y)
    finally:
        st.close()
