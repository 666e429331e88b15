"""BASELINE.json's FULL sizes, checked through size-independent properties (the strict oracle
cannot run 10^9 multiply-adds per query in test time):
  * two independent candidate generators -- the tcgen05 GEMM plan and the HBM scan plan, or the
    scan plan and the exhaustive fp64 plan -- must return bit-identical ids and distances;
  * every returned distance must equal the strict oracle's distance to that row, regenerated on
    the CPU from the counter-based corpus (bit-exact), results ascending, counts == k;
  * for one query per float config the ids are also checked against the chunked numpy fp64 tier
    over the whole corpus.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _winners_are_exact(oracle, st, slots, dists, qs, d, metric, rows_of=None, nq=3):
    for b in range(min(nq, qs.shape[0])):
        assert np.all(np.diff(dists[b]) >= 0)
        for j in range(slots.shape[1]):
            row = oracle.synth_f64(oracle.SEED_CORPUS, int(slots[b, j]), 1, d)[0] if rows_of is None \
                else rows_of(int(slots[b, j]))
            assert dists[b, j] == oracle.distance(qs[b], row, metric), (b, j)


def test_config2_1m_x_768_cosine_gemm_equals_scan_and_oracle(native, oracle):
    from erlvectordb_b200.device_store import DeviceStore
    n, d, k, B = 1_000_000, 768, 10, 96
    st = DeviceStore(dtype="f32", device=0)
    st.fill_synthetic(oracle.SEED_CORPUS, n, d)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, k, "cosine")
    assert st.stats()["last_plan"] == native.PLAN_GEMM and (gc == k).all()
    st.set_plan("scan")
    ss, sd, sc = st.search(qs[:12], k, "cosine")
    assert np.array_equal(gs[:12], ss) and np.array_equal(gd[:12], sd)
    _winners_are_exact(oracle, st, gs, gd, qs, d, "cosine")
    bi, bd = oracle.bulk_search_synth(oracle.SEED_CORPUS, n, d, qs[:1], k, "cosine")
    assert gs[:1].tolist() == bi.tolist()
    np.testing.assert_allclose(gd[:1], bd, rtol=1e-5, atol=0)   # north_star tolerance for fp32 stores
    st.close()


def test_config3_10m_x_128_euclidean_k100_gemm_equals_scan(native, oracle):
    from erlvectordb_b200.device_store import DeviceStore
    n, d, k, B = 10_000_000, 128, 100, 160
    st = DeviceStore(dtype="f32", device=0)
    st.fill_synthetic(oracle.SEED_CORPUS, n, d)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    st.set_plan("gemm")
    gs, gd, gc = st.search(qs, k, "euclidean")
    assert st.stats()["last_plan"] == native.PLAN_GEMM and (gc == k).all()
    st.set_plan("scan")
    ss, sd, sc = st.search(qs[:6], k, "euclidean")
    assert np.array_equal(gs[:6], ss) and np.array_equal(gd[:6], sd)
    _winners_are_exact(oracle, st, gs, gd, qs, d, "euclidean", nq=2)
    st.close()


def test_config4_shard_12m5_x_96_u8_scan_equals_exhaustive_plan(native, oracle):
    """One GPU's share of the 100M x 96 quantization_8bit store: the dp4a scan against the
    exhaustive fp64 plan (every row, stable radix sort)."""
    from erlvectordb_b200.device_store import DeviceStore
    n, d, k = 12_500_000, 96, 10
    st = DeviceStore(dtype="u8", device=0)
    st.fill_synthetic(oracle.SEED_CORPUS, n, d)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 2, d)
    ss, sd, sc = st.search(qs, k, "cosine")
    assert st.stats()["last_plan"] == native.PLAN_SCAN
    st.set_plan("exact")
    es, ed, ec = st.search(qs, k, "cosine")
    assert np.array_equal(ss, es) and np.array_equal(sd, ed)

    def deq(slot):
        codes, mn, sc_ = st.get_codes(slot)
        return oracle.dequantize_8bit(codes, mn, sc_)
    _winners_are_exact(oracle, st, ss, sd, qs, d, "cosine", rows_of=deq, nq=2)
    # the codes themselves: bit-identical to the reference codec on the regenerated rows
    for slot in (0, 4_999_999, n - 1):
        codes, mn, sc_ = st.get_codes(slot)
        c, omn, omx, osc = oracle.quantize_8bit(oracle.synth_f64(oracle.SEED_CORPUS, slot, 1, d)[0])
        assert codes.tolist() == list(c) and mn == omn and sc_ == osc
    st.close()


def test_config5_1m_x_1536_manhattan_and_4bit(native, oracle):
    from erlvectordb_b200.device_store import DeviceStore
    n, d, k = 1_000_000, 1536, 10
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 2, d)
    st = DeviceStore(dtype="f32", device=0)
    st.fill_synthetic(oracle.SEED_CORPUS, n, d)
    ss, sd, sc = st.search(qs, k, "manhattan")
    assert st.stats()["last_plan"] == native.PLAN_SCAN
    _winners_are_exact(oracle, st, ss, sd, qs, d, "manhattan", nq=2)
    bi, bd = oracle.bulk_search_synth(oracle.SEED_CORPUS, n, d, qs[:1], k, "manhattan")
    assert ss[:1].tolist() == bi.tolist()
    st.close()
    s4 = DeviceStore(dtype="u4", device=0)
    s4.fill_synthetic(oracle.SEED_CORPUS, n, d)
    a_s, a_d, a_c = s4.search(qs, k, "cosine")
    s4.set_plan("exact")
    e_s, e_d, e_c = s4.search(qs, k, "cosine")
    assert np.array_equal(a_s, e_s) and np.array_equal(a_d, e_d)
    s4.close()


def test_config5_manhattan_batch_shares_row_loads_and_stays_exact(native, oracle):
    """BASELINE configs[4] at full size, a BATCH: the multi-query scan (8 queries per pass over the 6 GB of
    rows) must return, for every query, exactly what the one-query pass returns, and the winners' distances
    are the strict oracle's."""
    from erlvectordb_b200.device_store import DeviceStore
    n, d, k, B = 1_000_000, 1536, 10, 11
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    st = DeviceStore(dtype="f32", device=0)
    st.fill_synthetic(oracle.SEED_CORPUS, n, d)
    bs, bd, bc = st.search(qs, k, "manhattan")             # 8 + 3 queries: a full and a ragged pass
    assert st.stats()["last_plan"] == native.PLAN_SCAN and (bc == k).all()
    for b in (0, 7, 8, 10):
        s1, d1, _ = st.search(qs[b], k, "manhattan")
        assert bs[b].tolist() == s1[0].tolist() and bd[b].tolist() == d1[0].tolist()
    _winners_are_exact(oracle, st, bs, bd, qs, d, "manhattan", nq=2)
    st.close()


def test_config2_one_handle_three_shards_equals_single_store(native, oracle):
    """BASELINE configs[1] at full size behind ONE multi-shard handle (three shards sharing this GPU): batch
    1024 and a lone query, under the default one-exchange scheme, bit-equal to the single-device store."""
    from erlvectordb_b200.device_store import DeviceStore
    n, d, k = 1_000_000, 768, 10
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 1024, d)
    one = DeviceStore(dtype="f32", device=0)
    one.fill_synthetic(oracle.SEED_CORPUS, n, d)
    want = one.search(qs, k, "cosine")
    w1 = one.search(qs[5], k, "cosine")
    one.close()
    m = DeviceStore(dtype="f32", devices=[0, 0, 0])
    m.fill_synthetic(oracle.SEED_CORPUS, n, d)
    got = m.search(qs, k, "cosine")
    g1 = m.search(qs[5], k, "cosine")
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    for a, b in zip(g1, w1):
        assert np.array_equal(a, b)
    _winners_are_exact(oracle, m, got[0], got[1], qs, d, "cosine", nq=2)
    m.close()


def test_config4_shard_query_batch_on_the_i8_plan(native, oracle):
    """One GPU's share of BASELINE configs[3] (12.5M x 96 quantization_8bit) answering a 256-query batch
    on the tcgen05 kind::i8 plan: every query equals the dp4a scan plan's answer, two of them the
    exhaustive fp64 plan's."""
    from erlvectordb_b200.device_store import DeviceStore
    n, d, k, B = 12_500_000, 96, 10, 256
    st = DeviceStore(dtype="u8", device=0)
    try:
        st.fill_synthetic(oracle.SEED_CORPUS, n, d)
        qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
        gs, gd, gc = st.search(qs, k, "cosine")
        assert st.stats()["last_plan"] == native.PLAN_GEMM and (gc == k).all()
        st.set_plan("scan")
        ss, sd, sc = st.search(qs, k, "cosine")
        assert st.stats()["last_plan"] == native.PLAN_SCAN
        assert np.array_equal(gs, ss) and np.array_equal(gd, sd)
        st.set_plan("exact")
        es, ed, ec = st.search(qs[:2], k, "cosine")
        assert np.array_equal(gs[:2], es) and np.array_equal(gd[:2], ed)
    finally:
        st.close()
