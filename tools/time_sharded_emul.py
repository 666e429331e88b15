"""Per-phase device time of the two-phase sharded search, G ranks emulated on ONE GPU (mailboxes
connected by pointer; each phase enqueued for all ranks before the next, so nothing waits).
    python tools/time_sharded_emul.py [G] [rows] [dim] [B] [k]
Prints the time of each phase group divided by G = what one rank pays."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from erlvectordb_b200 import synth
from erlvectordb_b200.device_store import (DeviceStore, Exchange, gemm_window, sharded_phase1, sharded_phase2,
                                           sharded_phase3)
from erlvectordb_b200.sharded import blob_words, shard_bounds

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 768
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
k = int(sys.argv[5]) if len(sys.argv) > 5 else 10
metric = sys.argv[6] if len(sys.argv) > 6 else "cosine"
dev = torch.device("cuda", 0)
kp = gemm_window(k, n)
stores, xws, xes = [], [], []
for g in range(G):
    lo, hi = shard_bounds(n, G, g)
    st = DeviceStore(dtype="f32", device=0)
    st.fill_synthetic(synth.SEED_CORPUS, hi - lo, d, row0=lo)
    stores.append((st, lo))
    xws.append(Exchange(0, g, G, B * kp + B))
    xes.append(Exchange(0, g, G, B * kp))
for xs in (xws, xes):
    boxes = [x.mailbox for x in xs]
    for x in xs:
        x.connect_ptrs(boxes)
q = torch.from_numpy(synth.synth(synth.SEED_QUERY, 0, B, d)).to(dev)
outs = [torch.zeros((blob_words(B, k),), dtype=torch.int64, device=dev) for _ in range(G)]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tot = [0.0, 0.0, 0.0]
iters = 6
for it in range(iters + 2):
    ev[0].record()
    for g, (st, lo) in enumerate(stores):
        assert sharded_phase1(st, xws[g], q.data_ptr(), B, d, k, metric, lo, n, 1) == 0
    ev[1].record()
    for g, (st, lo) in enumerate(stores):
        sharded_phase2(st, xws[g], xes[g], q.data_ptr(), B, k, metric, n, 1)
    ev[2].record()
    for g, (st, lo) in enumerate(stores):
        sharded_phase3(st, xes[g], B, k, metric, n, outs[g].data_ptr(), 1)
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        for i in range(3):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
print(f"G={G} rows={n} d={d} B={B} k={k} {metric}: per rank  phase1 (GEMM+window+push) {tot[0]/iters/G*1e3:.1f} us, "
      f"phase2 (merge+re-rank+push) {tot[1]/iters/G*1e3:.1f} us, phase3 (final) {tot[2]/iters/G*1e3:.1f} us")
