"""N-rank check of the row-sharded path on real GPUs (torchrun): every rank must end up with the
single-store result.  torchrun --nproc-per-node N tools/check_sharded.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from erlvectordb_b200 import synth
from erlvectordb_b200.device_store import DeviceStore
from erlvectordb_b200.sharded import ShardedStore

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for dtype, metric, n, d, B, k in [("f32", "cosine", 200_000, 128, 4, 10), ("f32", "cosine", 300_000, 256, 256, 10),
                                  ("f32", "euclidean", 200_000, 128, 128, 100), ("u8", "cosine", 400_000, 96, 1, 10)]:
    st = ShardedStore(dtype=dtype, device=local, rank=rank, world=world)
    st.fill_synthetic(synth.SEED_CORPUS, n, d)
    qh = synth.synth(synth.SEED_QUERY, 0, B, d)
    q = torch.from_numpy(qh).cuda()
    for _ in range(4):   # repeated searches: epoch/parity reuse of the peer mailboxes
        ids, dd, cnt, flags = st.search(q, k, metric)
    st2 = ShardedStore(dtype=dtype, device=local, rank=rank, world=world, exchange="nccl")
    st2.fill_synthetic(synth.SEED_CORPUS, n, d)
    i2, d2, c2, f2 = st2.search(q, k, metric)
    assert torch.equal(i2, ids) and torch.equal(d2, dd), "p2p and nccl exchanges disagree"
    st2.close()
    if rank == 0:
        print(f"  exchange used: {st.exchange}", flush=True)
    one = DeviceStore(dtype=dtype, device=local)
    one.fill_synthetic(synth.SEED_CORPUS, n, d)
    s_ids, s_d, s_c = one.search(qh, k, metric)
    same = np.array_equal(ids.cpu().numpy().astype(np.uint32), s_ids) and np.array_equal(dd.cpu().numpy(), s_d)
    ok &= same
    print(f"rank {rank}: {dtype} {metric} n={n} d={d} B={B} k={k}: {'identical' if same else 'MISMATCH'}", flush=True)
    st.close(); one.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_OK" if int(t.item()) == 1 else "SHARDED_MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
