"""Run under torchrun on N GPUs: the row-sharded search (NCCL allgather + merge kernel) must be
bit-identical to a single-GPU store holding the whole corpus.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from erlvectordb_b200 import synth
from erlvectordb_b200.sharded import ShardedStore

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (n, d, B, k, dtype, metric) in [(200_003, 768, 1, 10, "f32", "cosine"), (200_003, 768, 300, 10, "f32", "cosine"),
                                    (300_000, 96, 5, 10, "u8", "cosine"), (150_000, 128, 3, 100, "f32", "euclidean")]:
    sh = ShardedStore(dtype=dtype, device=local, rank=rank, world=world)
    sh.fill_synthetic(synth.SEED_CORPUS, n, d)
    q = torch.from_numpy(synth.synth(synth.SEED_QUERY, 0, B, d)).to(dev)
    ids, dd, cnt, flags = sh.search(q, k, metric)
    torch.cuda.synchronize()
    one = ShardedStore(dtype=dtype, device=local, rank=0, world=1)
    one.fill_synthetic(synth.SEED_CORPUS, n, d)
    ids1, dd1, cnt1, flags1 = one.search(q, k, metric)
    torch.cuda.synchronize()
    same = bool(torch.equal(ids, ids1) and torch.equal(dd, dd1) and torch.equal(cnt, cnt1))
    ok = ok and same and int(flags.sum()) == 0
    if rank == 0:
        print(f"n={n} d={d} B={B} k={k} {dtype} {metric}: sharded x{world} == single: {same}, flagged={int(flags.sum())}")
    sh.close(); one.close()
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("SHARDED_OK" if int(t.item()) == 1 else "SHARDED_MISMATCH")
sys.exit(0 if int(t.item()) == 1 else 1)
