"""BASELINE configs[3]: 100M x 96 quantization_8bit store row-sharded over the GPUs of one box,
int8 (dp4a) scan + exchange + merge, k = 10, query batch 1.  Weak scaling: 12.5M rows per GPU.

    python -m torch.distributed.run --nproc-per-node N tools/bench_cfg4.py [--rows-per-gpu 12500000]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from erlvectordb_b200 import synth
from erlvectordb_b200.sharded import ShardedStore

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=12_500_000)
ap.add_argument("--dim", type=int, default=96)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--warmup", type=int, default=20)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_total = a.rows_per_gpu * world
st = ShardedStore(dtype="u8", device=local, rank=rank, world=world)
st.fill_synthetic(synth.SEED_CORPUS, n_total, a.dim)
q = torch.from_numpy(synth.synth(synth.SEED_QUERY, 0, 1, a.dim)).to(dev)
for _ in range(a.warmup):
    out = st.search(q, 10, "cosine")
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
st._dev.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = st.search(q, 10, "cosine")
e1.record()
torch.cuda.synchronize()
ns, kms = st._dev.profile_read()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item()) / a.steps
if rank == 0:
    kern = kms / max(ns, 1)
    bytes_per_gpu = a.rows_per_gpu * (a.dim + 8)
    print(json.dumps({"case": f"cfg4 {n_total} x {a.dim} u8 cosine k=10 B=1, {world} GPU(s), {a.rows_per_gpu} rows per GPU (weak)",
                      "ms_per_query": round(ms, 4), "qps": round(1e3 / ms, 1), "rows_per_s": round(n_total / (ms * 1e-3), 1),
                      "scan_kernel_ms": round(kern, 4), "scan_gbs_per_gpu": round(bytes_per_gpu / (kern * 1e-3) / 1e9, 1),
                      "exchange": st.exchange, "top1": int(out[0][0, 0]), "flagged": int(out[3].sum())}))
st.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
