"""world_size-2 gloo test of the row-sharded search plumbing (bounds, allgather layout,
merge semantics) on CPU.  The CUDA local search / merge kernels are replaced by the oracle
here -- tests may call it; the product path (ShardedStore defaults) may not."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_everything():
    from erlvectordb_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000, 1_000_000):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0] and a[0] <= a[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, k, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    from erlvectordb_b200.sharded import ShardedStore, shard_bounds

    lo, hi = shard_bounds(n, world, rank)
    rows = O.synth_f64(O.SEED_CORPUS, lo, hi - lo, d) if hi > lo else np.zeros((0, d))

    q_all = torch.from_numpy(O.synth_f64(O.SEED_QUERY, 0, 3, d))
    calls = []

    def local_search(q, kk, metric, plan="auto", kp_min=0):
        # Stands in for the device path INCLUDING its contract on unproven windows: rank 1 cannot prove
        # query 1 on the first pass nor with the wider window (it returns a wrong, flagged result), only
        # the exhaustive plan settles it -- ShardedStore.search must climb the ladder on every rank.
        B = q.shape[0]
        calls.append((plan, kp_min, B))
        ids = torch.full((B, kk), -1, dtype=torch.int64)
        dd = torch.zeros((B, kk), dtype=torch.float64)
        cnt = torch.zeros((B,), dtype=torch.int32)
        flg = torch.zeros((B,), dtype=torch.int32)
        for b in range(B):
            if hi > lo:
                r, dist_ = O.search(rows, q[b].numpy(), kk, metric)
                unproven = rank == world - 1 and plan != "exact" and torch.equal(q[b], q_all[1])
                if unproven:
                    r, dist_ = r[::-1].copy(), dist_[::-1] + 1.0
                    flg[b] = 1
                ids[b, :len(r)] = torch.from_numpy(r + lo)
                dd[b, :len(r)] = torch.from_numpy(dist_)
                cnt[b] = len(r)
        return ids, dd, cnt, flg

    def merge(g_ids, g_d, g_c, kk):  # same contract as evdb_merge_topk_dev
        G, B = g_c.shape
        out_ids = torch.full((B, kk), -1, dtype=torch.int64)
        out_d = torch.zeros((B, kk), dtype=torch.float64)
        out_c = torch.zeros((B,), dtype=torch.int32)
        for b in range(B):
            cand = [(float(g_d[g, b, j]), int(g_ids[g, b, j])) for g in range(G) for j in range(int(g_c[g, b]))]
            cand.sort()
            cand = cand[:kk]
            out_c[b] = len(cand)
            for j, (x, i) in enumerate(cand):
                out_d[b, j], out_ids[b, j] = x, i
        return out_ids, out_d, out_c

    st = ShardedStore(rank=rank, world=world, local_search=local_search, merge=merge)
    st.fill_synthetic(O.SEED_CORPUS, n, d)
    assert (st.lo, st.hi) == (lo, hi)
    q = q_all
    ids, dd, cnt, flags = st.search(q, k, "cosine")
    assert int(flags.sum()) == 0 and st.n_escalations == 1
    # the ladder: everything once, then the flagged query alone with a 256-key scan window, then exact
    assert calls == [("auto", 0, 3), ("scan", 256, 1), ("exact", 0, 1)], calls
    raw = st.search(q, k, "cosine", escalate=False)
    assert raw[3].tolist() == [0, 1, 0]          # without the ladder the flag reaches the caller

    # replica group: every rank holds the whole store and answers a disjoint block of the queries
    from erlvectordb_b200.sharded import ReplicaGroup
    all_rows = O.synth_f64(O.SEED_CORPUS, 0, n, d)

    def replica_search(qb, kk, metric, plan="auto", kp_min=0):
        B = qb.shape[0]
        r_ids = torch.full((B, kk), -1, dtype=torch.int64)
        r_dd = torch.zeros((B, kk), dtype=torch.float64)
        r_cnt = torch.zeros((B,), dtype=torch.int32)
        r_flg = torch.zeros((B,), dtype=torch.int32)
        for b in range(B):
            r, dist_ = O.search(all_rows, qb[b].numpy(), kk, metric)
            if plan == "auto" and torch.equal(qb[b], q_all[2]):   # unproven on the first pass only
                r, dist_, r_flg[b] = r[::-1].copy(), dist_[::-1] + 1.0, 1
            r_ids[b, :len(r)] = torch.from_numpy(r)
            r_dd[b, :len(r)] = torch.from_numpy(dist_)
            r_cnt[b] = len(r)
        return r_ids, r_dd, r_cnt, r_flg

    rg = ReplicaGroup(rank=rank, world=world, local_search=replica_search)
    g_ids, g_dd, g_cnt, g_flg = rg.search(q, k, "cosine")
    assert torch.equal(g_ids, ids) and torch.equal(g_dd, dd) and torch.equal(g_cnt, cnt) and int(g_flg.sum()) == 0
    out_q.put((rank, ids.numpy(), dd.numpy(), cnt.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [600, 5])
def test_two_rank_sharded_search_equals_single_store(oracle, n):
    d, k, world = 32, 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, n, d)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 3, d)
    for rank, ids, dd, cnt in got:
        for b in range(3):
            r, dist_ = oracle.search(rows, qs[b], k, "cosine")
            assert cnt[b] == len(r)
            assert ids[b, :cnt[b]].tolist() == r.tolist()      # every rank holds the same global top-k
            assert dd[b, :cnt[b]].tolist() == dist_.tolist()   # bit-identical to the single-store result
