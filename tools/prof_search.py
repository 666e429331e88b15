"""Small driver for profiling one search configuration (used under ncu and for A/B timing).

    python tools/prof_search.py --rows 1000000 --dim 768 --batch 1024 --dtype f32 --metric cosine --iters 5
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from erlvectordb_b200 import synth
from erlvectordb_b200.sharded import ShardedStore

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--metric", default="cosine")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--plan", default="auto")
a = ap.parse_args()

torch.cuda.set_device(0)
st = ShardedStore(dtype=a.dtype, device=0, rank=0, world=1)
st.fill_synthetic(synth.SEED_CORPUS, a.rows, a.dim)
st._dev.set_plan(a.plan)
q = torch.from_numpy(synth.synth(synth.SEED_QUERY, 0, a.batch, a.dim)).cuda()
for _ in range(2):
    out = st.search(q, a.k, a.metric)
torch.cuda.synchronize()
st._dev.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    out = st.search(q, a.k, a.metric)
e1.record()
torch.cuda.synchronize()
n, ms = st._dev.profile_read()
stt = st._dev.stats()
print(f"rows={a.rows} dim={a.dim} batch={a.batch} dtype={a.dtype} metric={a.metric} plan={stt['last_plan']} "
      f"step_ms={e0.elapsed_time(e1)/a.iters:.4f} kernel_ms={ms/max(n,1):.4f} flagged={int(out[3].sum())} "
      f"debug={os.environ.get('EVDB_GEMM_DEBUG','0')}")
