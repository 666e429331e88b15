"""N-rank check of the row-sharded path on real GPUs: every rank must end up with the single-store
result, bit for bit, through the peer-memory exchange AND the NCCL cross-check, including queries
whose windows cannot be proven on the first pass (collective escalation).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded.py

Prints one line per case and rank, a digest of every result, and SHARDED_OK / SHARDED_MISMATCH.
(tests/test_gpu_multiproc.py runs it with 2 ranks when 2 GPUs are visible; its 8-GPU output is
kept under profiles/.)"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from erlvectordb_b200 import synth
from erlvectordb_b200.device_store import DeviceStore
from erlvectordb_b200.sharded import ReplicaGroup, ShardedStore, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True


def digest(ids, dd):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(ids, dtype=np.uint32).tobytes())
    h.update(np.ascontiguousarray(dd, dtype=np.float64).tobytes())
    return h.hexdigest()[:16]


CASES = [("f32", "cosine", 200_000, 128, 4, 10, False), ("f32", "cosine", 300_000, 256, 256, 10, False),
         ("f32", "euclidean", 200_000, 128, 128, 100, False), ("u8", "cosine", 400_000, 96, 1, 10, False),
         ("f32", "manhattan", 100_000, 64, 3, 10, False), ("u4", "cosine", 200_000, 64, 2, 10, False),
         ("f32", "cosine", 160_000, 64, 64, 10, True), ("f32", "euclidean", 160_000, 64, 3, 10, True),
         ("u8", "cosine", 400_000, 96, 64, 10, False)]   # quantization_8bit batch: tcgen05 kind::i8 plan on every shard
for dtype, metric, n, d, B, k, cluster in CASES:
    qh = synth.synth(synth.SEED_QUERY, 0, B, d)
    st = ShardedStore(dtype=dtype, device=local, rank=rank, world=world)
    one = DeviceStore(dtype=dtype, device=local)
    if cluster:
        # a cluster of near-duplicates of query 1: the candidate pass cannot separate them, the window
        # proof fails on the shards that hold them, every rank must climb the ladder together
        rows = synth.synth(synth.SEED_CORPUS, 0, n, d).astype(np.float32)
        rng = np.random.default_rng(5)
        for i in range(400):
            rows[11 + 97 * i] = (qh[1] * (1.0 + 2e-7 * rng.standard_normal(d))).astype(np.float32)
        lo, hi = shard_bounds(n, world, rank)
        st.bulk_load_shard(rows[lo:hi], n)
        one.bulk_load(rows)
    else:
        st.fill_synthetic(synth.SEED_CORPUS, n, d)
        one.fill_synthetic(synth.SEED_CORPUS, n, d)
    q = torch.from_numpy(qh).cuda()
    if B >= 16:          # tcgen05 batches: the window / owner two-phase scheme must agree with the default one-exchange scheme
        os.environ["EVDB_SHARD_TWO_PHASE"] = "1"
        t_ids, t_dd, _, t_fl = st.search(q, k, metric)
        t_ids, t_dd = t_ids.clone(), t_dd.clone()
        os.environ["EVDB_SHARD_TWO_PHASE"] = "0"
    for _ in range(4):   # repeated searches: epoch/parity reuse of the peer mailboxes
        ids, dd, cnt, flags = st.search(q, k, metric)
    if B >= 16:
        assert torch.equal(t_ids, ids) and torch.equal(t_dd, dd), "two-phase and one-exchange schemes disagree"
    assert int(flags.sum()) == 0, "an unproven result left ShardedStore.search"
    if not cluster:
        st2 = ShardedStore(dtype=dtype, device=local, rank=rank, world=world, exchange="nccl")
        st2.fill_synthetic(synth.SEED_CORPUS, n, d)
        i2, d2, c2, f2 = st2.search(q, k, metric)
        assert torch.equal(i2, ids) and torch.equal(d2, dd), "p2p and nccl exchanges disagree"
        st2.close()
    s_ids, s_d, s_c = one.search(qh, k, metric)
    same = np.array_equal(ids.cpu().numpy().astype(np.uint32), s_ids) and np.array_equal(dd.cpu().numpy(), s_d)
    if cluster and B >= 16:   # the tcgen05 pass cannot separate the cluster: the ladder must have been climbed
        same = same and st.n_escalations > 0
    ok &= same
    print(f"rank {rank}/{world}: {dtype} {metric} n={n} d={d} B={B} k={k}{' near-duplicate cluster' if cluster else ''}: "
          f"{'identical' if same else 'MISMATCH'} to the single store, digest {digest(ids.cpu().numpy(), dd.cpu().numpy())} "
          f"(single {digest(s_ids, s_d)}), exchange {st.exchange}, escalated {st.n_escalations}", flush=True)
    if not cluster and dtype == "f32" and B >= 4:
        rg = ReplicaGroup(dtype=dtype, device=local, rank=rank, world=world)
        rg.fill_synthetic(synth.SEED_CORPUS, n, d)
        r_ids, r_d, r_c, r_f = rg.search(q, k, metric)
        same = np.array_equal(r_ids.cpu().numpy().astype(np.uint32), s_ids) and np.array_equal(r_d.cpu().numpy(), s_d)
        ok &= same
        print(f"rank {rank}/{world}: replica group {metric} B={B}: {'identical' if same else 'MISMATCH'}", flush=True)
        rg.close()
    st.close(); one.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_OK" if int(t.item()) == 1 else "SHARDED_MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
