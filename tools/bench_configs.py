"""Device-timed runs of the other BASELINE.json configurations (parity-test cases, not bench lines):
reports step time, dominant-kernel time and the roofline fraction for each.

    python tools/bench_configs.py [--quick]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from erlvectordb_b200 import synth
from erlvectordb_b200.sharded import ShardedStore

PEAKS = {"hbm_gbs": 6454.9, "tf": 1414.3}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    j = json.load(open(p))
    PEAKS = {"hbm_gbs": j["hbm_gbs"], "tf": j.get("bf16_tflops_sustained", j["bf16_tflops"])}

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
a = ap.parse_args()
torch.cuda.set_device(0)

CASES = [
    # name, rows, dim, dtype, metric, k, batch, bytes_per_row
    ("cfg1 10k x128 f32 cosine k=10 B=1", 10_000, 128, "f32", "cosine", 10, 1, 128 * 4),
    ("cfg2 1M x768 f32 cosine k=10 B=1", 1_000_000, 768, "f32", "cosine", 10, 1, 768 * 4),
    ("cfg2 1M x768 f32 cosine k=10 B=8", 1_000_000, 768, "f32", "cosine", 10, 8, 768 * 4),
    ("cfg2 1M x768 f32 cosine k=10 B=1024", 1_000_000, 768, "f32", "cosine", 10, 1024, 768 * 4),
    ("cfg3 10M x128 f32 euclidean k=100 B=1", 10_000_000, 128, "f32", "euclidean", 100, 1, 128 * 4),
    ("cfg3 10M x128 f32 euclidean k=100 B=4096 (tcgen05 GEMM + top-k epilogue)", 10_000_000, 128, "f32", "euclidean", 100, 4096, 128 * 4),
    ("cfg3' 10M x128 f32 cosine k=10 B=4096", 10_000_000, 128, "f32", "cosine", 10, 4096, 128 * 4),
    ("cfg4/8 12.5M x96 u8 cosine k=10 B=1", 12_500_000, 96, "u8", "cosine", 10, 1, 96 + 8),
    ("cfg4' 12.5M x128 u8 cosine k=10 B=1", 12_500_000, 128, "u8", "cosine", 10, 1, 128 + 8),
    ("cfg5 1M x1536 f32 manhattan k=10 B=1", 1_000_000, 1536, "f32", "manhattan", 10, 1, 1536 * 4),
    ("cfg5 1M x1536 u4 cosine k=10 B=1", 1_000_000, 1536, "u4", "cosine", 10, 1, 1536 // 2 + 8),
    ("extra 1M x768 bf16 cosine k=10 B=1", 1_000_000, 768, "bf16", "cosine", 10, 1, 768 * 2),
]
out = []
for name, n, d, dtype, metric, k, B, bpr in CASES:
    if a.quick and n > 1_000_000:
        continue
    st = ShardedStore(dtype=dtype, device=0, rank=0, world=1)
    st.fill_synthetic(synth.SEED_CORPUS, n, d)
    q = torch.from_numpy(synth.synth(synth.SEED_QUERY, 0, B, d)).cuda()
    iters = 30 if B == 1 else 5
    for _ in range(3):
        o = st.search(q, k, metric)
    torch.cuda.synchronize()
    st._dev.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        o = st.search(q, k, metric)
    e1.record()
    torch.cuda.synchronize()
    ns, kms = st._dev.profile_read()
    stt = st._dev.stats()
    step = e0.elapsed_time(e1) / iters
    kern = kms / max(ns, 1)
    rec = {"case": name, "plan": {1: "scan", 2: "gemm", 3: "exact"}.get(stt["last_plan"]), "step_ms": round(step, 4),
           "kernel_ms": round(kern, 4), "qps": round(B / step * 1e3, 1), "flagged": int(o[3].sum())}
    if stt["last_plan"] == 2 and B < 128:
        # few queries: the tcgen05 candidate pass is bound by reading its 2-byte operand column once
        col = float(n) * ((d + 63) // 64 * 64) * 2
        gbs = col / (kern * 1e-3) / 1e9
        rec.update({"bound": "hbm", "achieved_gbs": round(gbs, 1), "frac": round(gbs / PEAKS["hbm_gbs"], 3),
                    "bytes": "fp16 operand column, read once for the batch",
                    "fp32_rows_equivalent_gbs": round(float(n) * bpr * B / (kern * 1e-3) / 1e9, 1)})
    elif stt["last_plan"] == 2:
        tf = 2.0 * n * d * B / (kern * 1e-3) / 1e12
        rec.update({"bound": "tensor", "achieved_tflops": round(tf, 1), "frac": round(tf / PEAKS["tf"], 3)})
    elif stt["last_plan"] == 1:
        gbs = float(n) * bpr * B / (kern * 1e-3) / 1e9
        rec.update({"bound": "hbm", "achieved_gbs": round(gbs, 1), "frac": round(gbs / PEAKS["hbm_gbs"], 3)})
    print(json.dumps(rec), flush=True)
    out.append(rec)
    st.close()
    del st
    torch.cuda.empty_cache()
