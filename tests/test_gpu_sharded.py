"""Row-sharded search on ONE GPU: G shards as G store handles, each writing its result into a packed
blob (evdb_store_search_dev), blobs laid out as the NCCL allgather would leave them, merged by
evdb_merge_topk_packed_dev.  Must equal the single-store result bit for bit (ids and fp64 distances);
the cross-process exchange itself is covered on CPU by tests/test_sharded_gloo.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,metric,n,d,B,k,G", [
    ("f32", "cosine", 40_000, 96, 3, 10, 4),       # scan plan per shard
    ("f32", "cosine", 90_000, 128, 160, 10, 3),    # tcgen05 GEMM plan per shard, ragged shard sizes
    ("f32", "euclidean", 60_000, 64, 64, 100, 2),  # euclidean GEMM plan, k = 100
    ("u8", "cosine", 50_000, 96, 2, 10, 8),        # BASELINE configs[3] shape scaled down: int8 scan, 8 shards
    ("f32", "cosine", 5, 16, 2, 10, 4),            # more shards than rows per shard: empty shards, k > N
    ("f32", "cosine", 700_000, 128, 2, 10, 2),     # two queries, shards past the AUTO crossover: tcgen05 plan per shard
])
def test_sharded_equals_single_store(native, oracle, dtype, metric, n, d, B, k, G):
    import torch
    from erlvectordb_b200.device_store import DeviceStore, merge_topk_packed_dev
    from erlvectordb_b200.sharded import blob_views, blob_words, shard_bounds

    dev = torch.device("cuda", 0)
    q = torch.from_numpy(oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)).to(dev)
    w = blob_words(B, k)
    gathered = torch.zeros((G, w), dtype=torch.int64, device=dev)
    stores = []
    for g in range(G):
        lo, hi = shard_bounds(n, G, g)
        if hi == lo:
            continue
        st = DeviceStore(dtype=dtype, device=0)
        st.fill_synthetic(oracle.SEED_CORPUS, hi - lo, d, row0=lo)
        ids, dists, counts, flags = blob_views(gathered[g], B, k)
        st.search_dev(q.data_ptr(), B, d, k, metric, lo, ids.data_ptr(), dists.data_ptr(), counts.data_ptr(),
                      flags.data_ptr(), 1)
        stores.append(st)
    merged = torch.zeros((w,), dtype=torch.int64, device=dev)
    merge_topk_packed_dev(0, gathered.data_ptr(), G, B, k, merged.data_ptr(), 1)
    torch.cuda.synchronize()
    m_ids, m_d, m_c, m_f = [t.cpu().numpy() for t in blob_views(merged, B, k)]
    assert int(m_f.sum()) == 0

    one = DeviceStore(dtype=dtype, device=0)
    one.fill_synthetic(oracle.SEED_CORPUS, n, d)
    s_ids, s_d, s_c = one.search(q.cpu().numpy(), k, metric)
    for b in range(B):
        c = int(s_c[b])
        assert int(m_c[b]) == c == min(k, n)
        assert m_ids[b, :c].tolist() == s_ids[b, :c].tolist()
        assert m_d[b, :c].tolist() == s_d[b, :c].tolist()
    for st in stores:
        st.close()
    one.close()


def test_peer_memory_exchange_equals_packed_merge(native, oracle):
    """evdb_exchange_*: G "ranks" of one process on one GPU (mailboxes connected by pointer instead of
    CUDA IPC).  All pushes are enqueued before any merge, so no kernel ever waits here; several
    searches in a row exercise the epoch/parity double buffering.  Every rank's merged blob must
    equal evdb_merge_topk_packed_dev over the same inputs."""
    import torch
    from erlvectordb_b200.device_store import DeviceStore, Exchange, merge_topk_packed_dev
    from erlvectordb_b200.sharded import blob_views, blob_words, shard_bounds

    G, n, d, B, k = 3, 30_000, 64, 5, 10
    dev = torch.device("cuda", 0)
    w = blob_words(B, k)
    stores = []
    for g in range(G):
        lo, hi = shard_bounds(n, G, g)
        st = DeviceStore(dtype="f32", device=0)
        st.fill_synthetic(oracle.SEED_CORPUS, hi - lo, d, row0=lo)
        stores.append((st, lo))
    xs = [Exchange(0, g, G, w) for g in range(G)]
    boxes = [x.mailbox for x in xs]
    for x in xs:
        x.connect_ptrs(boxes)
    for step in range(5):   # > 2 searches: both parities reused
        q = torch.from_numpy(oracle.synth_f64(oracle.SEED_QUERY, step * B, B, d)).to(dev)
        gathered = torch.zeros((G, w), dtype=torch.int64, device=dev)
        for g, (st, lo) in enumerate(stores):
            ids, dists, counts, flags = blob_views(gathered[g], B, k)
            st.search_dev(q.data_ptr(), B, d, k, "cosine", lo, ids.data_ptr(), dists.data_ptr(), counts.data_ptr(),
                          flags.data_ptr(), 1)
        ref = torch.zeros((w,), dtype=torch.int64, device=dev)
        merge_topk_packed_dev(0, gathered.data_ptr(), G, B, k, ref.data_ptr(), 1)
        for g in range(G):
            xs[g].push(gathered[g].data_ptr(), B, k, 1)
        outs = [torch.zeros((w,), dtype=torch.int64, device=dev) for _ in range(G)]
        for g in range(G):
            xs[g].merge(B, k, outs[g].data_ptr(), 1)
        torch.cuda.synchronize()
        for g in range(G):
            assert torch.equal(outs[g], ref), (step, g)
    for x in xs:
        x.close()
    for st, _ in stores:
        st.close()


@pytest.mark.parametrize("metric,n,d,B,k,G", [
    ("cosine", 90_000, 128, 160, 10, 4),
    ("cosine", 40_000, 768, 64, 10, 8),       # the bench shape, 8 shards
    ("euclidean", 60_000, 64, 33, 100, 2),    # 128-key windows, ragged batch
    ("cosine", 2_100, 32, 16, 10, 8),         # shards of ~262 rows: windows cover whole shards
])
def test_two_phase_sharded_search_equals_single_store(native, oracle, metric, n, d, B, k, G):
    """Windows travel, owners re-rank (evdb_store_search_sharded_phase1/2/3): G ranks emulated on one
    GPU with pointer-connected mailboxes; each phase is enqueued for ALL ranks before the next, so
    no kernel ever waits.  Every rank's packed result must equal the single-store search bit for bit."""
    import torch
    from erlvectordb_b200.device_store import (DeviceStore, Exchange, gemm_window, sharded_phase1, sharded_phase2,
                                               sharded_phase3)
    from erlvectordb_b200.sharded import blob_views, blob_words, shard_bounds

    dev = torch.device("cuda", 0)
    kp = gemm_window(k, n)
    stores, xws, xes = [], [], []
    for g in range(G):
        lo, hi = shard_bounds(n, G, g)
        st = DeviceStore(dtype="f32", device=0)
        st.fill_synthetic(oracle.SEED_CORPUS, hi - lo, d, row0=lo)
        stores.append((st, lo))
        xws.append(Exchange(0, g, G, B * kp + B))
        xes.append(Exchange(0, g, G, B * kp))
    for xs in (xws, xes):
        boxes = [x.mailbox for x in xs]
        for x in xs:
            x.connect_ptrs(boxes)
    one = DeviceStore(dtype="f32", device=0)
    one.fill_synthetic(oracle.SEED_CORPUS, n, d)
    for step in range(3):
        qh = oracle.synth_f64(oracle.SEED_QUERY, step * B, B, d)
        q = torch.from_numpy(qh).to(dev)
        outs = [torch.zeros((blob_words(B, k),), dtype=torch.int64, device=dev) for _ in range(G)]
        for g, (st, lo) in enumerate(stores):
            assert sharded_phase1(st, xws[g], q.data_ptr(), B, d, k, metric, lo, n, 1) == 0
        for g, (st, lo) in enumerate(stores):
            sharded_phase2(st, xws[g], xes[g], q.data_ptr(), B, k, metric, n, 1)
        for g, (st, lo) in enumerate(stores):
            sharded_phase3(st, xes[g], B, k, metric, n, outs[g].data_ptr(), 1)
        torch.cuda.synchronize()
        s_ids, s_d, s_c = one.search(qh, k, metric)
        for g in range(G):
            ids, dd, cnt, flags = [t.cpu().numpy() for t in blob_views(outs[g], B, k)]
            assert int(flags.sum()) == 0, (step, g)
            assert np.array_equal(cnt, s_c)
            assert np.array_equal(ids.astype(np.uint32), s_ids), (step, g)
            assert np.array_equal(dd, s_d), (step, g)
    for x in xws + xes:
        x.close()
    for st, _ in stores:
        st.close()
    one.close()


@pytest.mark.parametrize("metric,B", [("cosine", 40), ("manhattan", 5)])
def test_replica_group_and_sharded_store_classes_on_one_gpu(native, oracle, metric, B):
    """SURVEY 8f-4 (cluster_manager.erl:148-171: whole-store replicas) and the row-sharded class with a world
    of one: the CUDA paths behind ReplicaGroup.search / ShardedStore.search (escalating API, device-resident
    queries) return what the host entry point returns for the same store -- including a near-duplicate
    cluster that the first pass cannot prove."""
    import torch
    from erlvectordb_b200.device_store import DeviceStore
    from erlvectordb_b200.sharded import ReplicaGroup, ShardedStore
    n, d, k = 30_000, 64, 10
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d).astype(np.float32)
    qh = oracle.synth_f64(oracle.SEED_QUERY, 0, B, d)
    rng = np.random.default_rng(9)
    for i in range(300):
        rows[5 + 61 * i] = (qh[2] * (1.0 + 2e-7 * rng.standard_normal(d))).astype(np.float32)
    one = DeviceStore(dtype="f32", device=0)
    one.bulk_load(rows)
    want = one.search(qh, k, metric)
    q = torch.from_numpy(qh).cuda()
    rg = ReplicaGroup(dtype="f32", device=0, rank=0, world=1)
    rg.bulk_load(rows)
    ss = ShardedStore(dtype="f32", device=0, rank=0, world=1)
    ss.bulk_load_shard(rows, n)
    try:
        for st in (rg, ss):
            ids, dd, cnt, flags = st.search(q, k, metric)
            assert int(flags.sum()) == 0
            assert np.array_equal(ids.cpu().numpy().astype(np.uint32), want[0])
            assert np.array_equal(dd.cpu().numpy(), want[1])
        if metric == "cosine":
            assert rg.n_escalations + ss.n_escalations >= 2      # the cluster forced the ladder in both classes
    finally:
        rg.close(); ss.close(); one.close()
