"""Device-timed sweep over ad-hoc search configurations (tuning aid, not a bench line).

    python tools/sweep.py rows,dim,dtype,metric,k,batch[,plan[,iters]] ...

Prints one JSON line per case: step time, dominant-kernel time (library CUDA events), the
roofline figure that bounds the plan that ran, and how many queries were flagged.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from erlvectordb_b200 import synth
from erlvectordb_b200.sharded import ShardedStore

BPR = {"f32": lambda d: 4 * d, "bf16": lambda d: 2 * d, "u8": lambda d: d + 8, "u4": lambda d: (d + 1) // 2 + 8}
torch.cuda.set_device(0)
cache = {}
for spec in sys.argv[1:]:
    f = spec.split(",")
    n, d, dtype, metric, k, B = int(f[0]), int(f[1]), f[2], f[3], int(f[4]), int(f[5])
    plan = f[6] if len(f) > 6 else "auto"
    iters = int(f[7]) if len(f) > 7 else (30 if B == 1 else 5)
    key = (n, d, dtype)
    if key not in cache:
        for st in cache.values():
            st.close()
        cache.clear()
        torch.cuda.empty_cache()
        st = ShardedStore(dtype=dtype, device=0, rank=0, world=1)
        st.fill_synthetic(synth.SEED_CORPUS, n, d)
        cache[key] = st
    st = cache[key]
    st._dev.set_plan(plan)
    q = torch.from_numpy(synth.synth(synth.SEED_QUERY, 0, B, d)).cuda()
    for _ in range(3):
        o = st.search(q, k, metric, escalate=False)
    torch.cuda.synchronize()
    noprof = os.environ.get("SWEEP_NOPROF") == "1"   # no event brackets between the kernels of a search (PDL A/B)
    if not noprof:
        st._dev.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    e0.record()
    t0 = time.perf_counter()
    for _ in range(iters):
        o = st.search(q, k, metric, escalate=False)
    host_us = (time.perf_counter() - t0) / iters * 1e6   # host time to ENQUEUE one search (no sync)
    e1.record()
    torch.cuda.synchronize()
    ns, kms = (0, 0.0) if noprof else st._dev.profile_read()
    st._dev.profile(False)
    stt = st._dev.stats()
    step = e0.elapsed_time(e1) / iters
    kern = kms / max(ns, 1)
    rec = {"case": spec, "plan": {1: "scan", 2: "gemm", 3: "exact"}.get(stt["last_plan"]),
           "step_ms": round(step, 4), "kernel_ms": round(kern, 4), "qps": round(B / step * 1e3, 1),
           "flagged": int(o[3].sum()), "host_enqueue_us": round(host_us, 1)}
    if stt["last_plan"] == 2 and kern > 0:
        rec["tflops"] = round(2.0 * n * d * B / (kern * 1e-3) / 1e12, 1)
    elif stt["last_plan"] == 1 and kern > 0:
        rec["gbs"] = round(float(n) * BPR[dtype](d) * B / (kern * 1e-3) / 1e9, 1)
    print(json.dumps(rec), flush=True)
