"""ONE handle, several devices (evdb_opts.n_shards): the multi-device store must answer exactly as
the single-device store does on the same rows -- same slots, bit-equal fp64 distances -- through
every host entry point of the C ABI (reference seam: src/vector_store.erl:113-190).

On a one-GPU box the shards share device 0 (ordinals may repeat: the phases are then enqueued in
lock-step on one stream); with >= 2 visible GPUs the same tests also run on distinct devices, one
host thread + stream per device, candidates meeting over peer memory.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    sets = [[0, 0, 0], [0, 0]]
    n = torch.cuda.device_count()
    if n >= 2:
        sets.append(list(range(min(n, 8))))
        sets.append([0, 1])
    return sets


@pytest.fixture(params=range(4), ids=["3-on-one-gpu", "2-on-one-gpu", "all-gpus", "2-gpus"])
def devices(request):
    sets = _device_sets()
    if request.param >= len(sets):
        pytest.skip("needs >= 2 visible GPUs")
    return sets[request.param]


def _pair(native, devices, dtype="f32", **kw):
    from erlvectordb_b200.device_store import DeviceStore
    return DeviceStore(dtype=dtype, devices=devices, **kw), DeviceStore(dtype=dtype, device=0, **kw)


def _same(a, b):
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def _rows(oracle, n, d, seed_off=0):
    return oracle.synth_f64(oracle.SEED_CORPUS, seed_off, n, d)


@pytest.mark.parametrize("metric", ["cosine", "euclidean", "manhattan"])
def test_bulk_load_and_lone_queries_equal_single_store(native, oracle, devices, metric):
    n, d, k = 5003, 64, 10
    rows = _rows(oracle, n, d)
    m, s = _pair(native, devices)
    try:
        m.bulk_load(rows); s.bulk_load(rows)
        assert m.stats()["count"] == n and m.stats()["n_shards"] == len(devices)
        qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 5, d)
        got, want = m.search(qs, k, metric), s.search(qs, k, metric)
        _same(got, want)
        for b in range(5):      # and both equal the oracle
            r, dd = oracle.search(rows, qs[b], k, metric)
            assert got[0][b].tolist() == r.tolist() and got[1][b].tolist() == dd.tolist()
        _same(m.search(qs[0], k, metric), s.search(qs[0], k, metric))       # B = 1 < shards: empty slices
        got = m.search(qs, n + 5, metric)                                    # K > N -> all N rows, in order
        want = s.search(qs, n + 5, metric)
        _same(got, want)
    finally:
        m.close(); s.close()


@pytest.mark.parametrize("scheme", ["one-exchange", "two-phase"])
@pytest.mark.parametrize("metric,k,B", [("cosine", 10, 64), ("euclidean", 100, 130), ("cosine", 10, 1024)])
def test_tcgen05_batches_two_phase_equal_single_store(native, oracle, devices, metric, k, B, scheme, monkeypatch):
    """tcgen05 batches inside the handle, under both cross-shard schemes: finished per-shard results travel
    once (the default), or approximate windows travel and owners re-rank (EVDB_SHARD_TWO_PHASE=1)."""
    monkeypatch.setenv("EVDB_SHARD_TWO_PHASE", "1" if scheme == "two-phase" else "0")
    n, d = 6000 * len(devices), 128
    m, s = _pair(native, devices)
    try:
        m.fill_synthetic(oracle.SEED_CORPUS, n, d); s.fill_synthetic(oracle.SEED_CORPUS, n, d)
        qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
        for _ in range(3):   # consecutive searches reuse the mailboxes (epoch parity)
            got = m.search(qs, k, metric)
        assert m.stats()["last_plan"] == native.PLAN_GEMM
        _same(got, s.search(qs, k, metric))
        f32 = m.search(qs.astype(np.float32), k, metric)                     # fp32 host queries widen on the device
        _same(f32, got)
    finally:
        m.close(); s.close()


def test_insert_delete_stream_equals_single_store(native, oracle, devices):
    """The gen_server's op stream (insert new / overwrite / delete = swap with last), one op at a time:
    after every step both stores hold the same rows in the same slots and answer alike."""
    d, k = 24, 6
    rng = np.random.default_rng(7)
    pool = _rows(oracle, 400, d)
    m, s = _pair(native, devices)
    try:
        count = 0
        q = oracle.synth_f64(oracle.SEED_QUERY, 0, 2, d)
        for step in range(160):
            op = rng.integers(0, 10)
            if count == 0 or op < 5:                      # append
                v = pool[rng.integers(0, 400)]
                assert m.upsert(count, v) == 0 and s.upsert(count, v) == 0
                count += 1
            elif op < 7:                                  # overwrite
                slot = int(rng.integers(0, count))
                v = pool[rng.integers(0, 400)]
                assert m.upsert(slot, v) == 0 and s.upsert(slot, v) == 0
            else:                                         # delete
                slot = int(rng.integers(0, count))
                assert m.delete(slot) == s.delete(slot)
                count -= 1
            assert m.stats()["count"] == count == s.stats()["count"]
            if step % 8 == 0 or count < 4:
                if count:
                    _same(m.search(q, k, "cosine"), s.search(q, k, "cosine"))
                    probe = int(rng.integers(0, count))
                    assert np.array_equal(m.get(probe), s.get(probe))
                else:
                    assert m.search(q, k, "cosine")[2].tolist() == [0, 0]
        first_m, first_s = m.append(pool[:57]), s.append(pool[:57])      # batched append crosses every shard
        assert first_m == first_s == count
        _same(m.search(q, k, "euclidean"), s.search(q, k, "euclidean"))
        for slot in range(count + 57):
            assert np.array_equal(m.get(slot), s.get(slot))
        st = m.stats()
        assert st["upserts"] > 0 and st["deletes"] > 0
    finally:
        m.close(); s.close()


@pytest.mark.parametrize("dtype", ["u8", "u4"])
def test_quantized_store_and_compressed_records(native, oracle, devices, dtype):
    n, d, k = 3001, 48, 10
    rows = _rows(oracle, n, d)
    m, s = _pair(native, devices, dtype=dtype)
    try:
        m.bulk_load(rows); s.bulk_load(rows)
        qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 4, d)
        _same(m.search(qs, k, "cosine"), s.search(qs, k, "cosine"))
        _same(m.search(qs, k, "manhattan"), s.search(qs, k, "manhattan"))   # exhaustive plan on codes
        for slot in (0, 1, 2, 3, n - 1):
            cm, cs = m.get_codes(slot), s.get_codes(slot)
            assert np.array_equal(cm[0], cs[0]) and cm[1:] == cs[1:]
        # compressed records straight to the device columns (vector_persistence:load_vectors)
        codes = np.stack([s.get_codes(i)[0] for i in range(200)])
        mins = np.array([s.get_codes(i)[1] for i in range(200)])
        scales = np.array([s.get_codes(i)[2] for i in range(200)])
        m.bulk_load_codes(codes, mins, scales, d); s.bulk_load_codes(codes, mins, scales, d)
        assert m.stats()["count"] == 200
        _same(m.search(qs, k, "cosine"), s.search(qs, k, "cosine"))
    finally:
        m.close(); s.close()


def test_unprovable_windows_climb_the_ladder_on_every_shard(native, oracle, devices):
    """A cluster of near-duplicates of the query: the fp16 candidate pass cannot separate them, the
    window proof fails, and the handle must re-issue those queries (wider scan window, then the
    exhaustive plan) -- the caller still gets the exact answer."""
    n, d, k, B = 4000 * len(devices), 64, 10, 32
    rows = _rows(oracle, n, d).astype(np.float32)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    rng = np.random.default_rng(3)
    for i in range(300):   # 300 rows within ~1e-7 relative of query 5
        rows[17 + 13 * i] = (qs[5] * (1.0 + 2e-7 * rng.standard_normal(d))).astype(np.float32)
    m, s = _pair(native, devices)
    try:
        m.bulk_load(rows); s.bulk_load(rows)
        for metric in ("cosine", "euclidean"):
            got, want = m.search(qs, k, metric), s.search(qs, k, metric)
            _same(got, want)
            r, dd = oracle.search(rows.astype(np.float64), qs[5], k, metric)
            assert got[0][5].tolist() == r.tolist() and got[1][5].tolist() == dd.tolist()
        assert m.stats()["escalations"] >= 1
    finally:
        m.close(); s.close()


def test_reference_error_behaviour(native, oracle, devices):
    from erlvectordb_b200.device_store import DeviceStore
    m = DeviceStore(dtype="f32", devices=devices)
    try:
        q = np.ones((2, 5))
        assert m.search(q, 3)[2].tolist() == [0, 0]                    # empty store: {ok, []} for any length
        bad = np.ones(5); bad[3] = np.nan
        assert m.upsert(0, bad) == native.E_BAD_VECTOR                 # rejected ...
        assert m.upsert(0, np.ones(7)) == 0                            # ... and did not pin the dimension
        assert m.upsert(1, np.ones(5)) == native.E_DIM_MISMATCH
        assert m.search(np.ones((1, 5)), 3) == native.E_DIM_MISMATCH
        rows = _rows(oracle, 50, 7)
        m.bulk_load(rows)
        qs = np.ones((6, 7)); qs[5, 2] = np.inf                        # the bad element sits in the last shard's slice
        assert m.search(qs, 3) == native.E_BAD_VECTOR
        assert m.search(np.ones((6, 7)), 0)[2].tolist() == [0] * 6
        got = m.search(np.ones((6, 7)), 3)
        r, dd = oracle.search(rows, np.ones(7), 3, "cosine")
        assert got[0][4].tolist() == r.tolist() and got[1][4].tolist() == dd.tolist()
    finally:
        m.close()
