#!/usr/bin/env python
"""bench.py -- QPS of the search hot path on the BASELINE.json workload.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): 1,000,000 x 768 fp32 synthetic corpus, cosine, k = 10,
query batch 1024 (the headline `value`, tensor-bound) and batch 1 (reported under "batch1",
HBM-bound).  A step is one search call over one batch.  N > 1 row-shards the SAME corpus over
the ranks (strong scaling): each GPU searches its shard, candidates are all-gathered over
NCCL and merged on every rank.

`value`   : device-timed (CUDA events), queries already resident in HBM.
`e2e`     : the same metric through the reference-facing call with HOST buffers -- the C-ABI
            entry evdb_store_search_f64 at N = 1 (ShardedStore.search + copies at N > 1) -- with
            the host->device query copy and the device->host result copy inside the timed region.
`roofline`: dominant kernel (scan: HBM bytes; tcgen05 GEMM: flops) timed with CUDA events on its
            launching stream inside the timed region, against MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (a port of the reference's Erlang arithmetic: full fp64 scan with
            the query norm recomputed per row + full sort) on a bounded sample, rank 0, N = 1.
The corpus (3.07 GB; 1.5 GB bf16 shadow) is far larger than the 126 MB L2, so consecutive steps
cannot be served from cache ("inputs larger than L2").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, DIM, K = 1_000_000, 768, 10
BATCH_MAIN, BATCH_ONE = 1024, 1
METRIC_NAME = "QPS (k=10, 1Mx768 cosine, batch 1024)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"],
                "tf_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# reference arm: the reference's algorithm (CPU oracle port) on the host cores
# ---------------------------------------------------------------------------------------
def cpu_reference_qps(threads: int, sample_rows: int, steps: int, warmup: int):
    import numpy as np
    from oracle import oracle as O
    rows = O.synth_f32(O.SEED_CORPUS, 0, sample_rows, DIM, threads=threads)
    qs = O.synth_f64(O.SEED_QUERY, 0, threads * (steps + warmup), DIM)
    times = []
    for s in range(steps + warmup):
        q = qs[s * threads:(s + 1) * threads]
        t0 = time.perf_counter()
        O.search_f32_replicas(rows, q, K, "cosine")
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    # T replicas answer T queries per step over sample_rows rows; a full-corpus query costs
    # N_ROWS / sample_rows times that (linear scan; the N log N sort only makes it worse)
    qps = threads / t * (sample_rows / N_ROWS)
    return qps, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_rows = 100_000
    qps, t = cpu_reference_qps(threads, sample_rows, args.steps, args.warmup)
    sample = (f"{threads} threads x 1 query each per step over the first {sample_rows} rows of the "
              f"1Mx768 corpus (full fp64 scan, query norm recomputed per row, full sort); "
              f"QPS scaled by {sample_rows}/{N_ROWS}")
    out = {
        "impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "1Mx768 fp32 cosine k=10 (BASELINE.json configs[1])", "batch": threads,
                   "note": "reference algorithm (Erlang vector_store:perform_search/3) as a compiled C "
                           "port; no Erlang/OTP toolchain in the image, a BEAM run would be slower"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from erlvectordb_b200 import _native as N
    from erlvectordb_b200 import synth
    from erlvectordb_b200.sharded import ShardedStore

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; erlvectordb_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rc = N.lib().evdb_init(None, 0)
    if rc != 0:
        raise SystemExit(f"evdb_init: {N.lib().evdb_strerror(rc).decode()}")

    st = ShardedStore(dtype="f32", device=local, rank=rank, world=world)
    st.fill_synthetic(synth.SEED_CORPUS, N_ROWS, DIM)
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batch: int, steps: int, warmup: int, sample_clocks: bool):
        qh = torch.from_numpy(synth.synth(synth.SEED_QUERY, 0, batch, DIM)).pin_memory()
        qd = qh.to(dev)
        for _ in range(warmup):
            out = st.search(qd, K, "cosine")
        barrier()
        st._dev.profile(True)
        l0 = st._dev.stats()["kernel_launches"]
        sampler = ClockSampler(local)
        if sample_clocks:
            # nvidia-smi samples every 200 ms and the timed region may last only tens of ms: the
            # sampler also covers ~0.6 s of the SAME search loop run (untimed) right before it, so
            # that the clocks/throttle record describes this load, and the timed steps start from
            # the steady (power-capped) state rather than from a cold burst
            sampler.start()
            t_probe = time.perf_counter()
            while time.perf_counter() - t_probe < 0.6:
                for _ in range(8):
                    st.search(qd, K, "cosine")
                torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            out = st.search(qd, K, "cosine")
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sample_clocks else None
        nsamp, kms = st._dev.profile_read()
        st._dev.profile(False)
        launches = st._dev.stats()["kernel_launches"] - l0 + (steps if world > 1 else 0)  # + merge kernel
        plan = st._dev.stats()["last_plan"]
        flagged = int(out[3].sum().item())
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

        # ---- end to end: host buffers in, host results out, copies inside the timed region ----
        h_ids = torch.empty((batch, K), dtype=torch.int64).pin_memory()
        h_d = torch.empty((batch, K), dtype=torch.float64).pin_memory()
        if world == 1:
            import ctypes as C
            qn = qh.numpy()
            slots = np.empty((batch, K), dtype=np.uint32)
            dists = np.empty((batch, K), dtype=np.float64)
            counts = np.empty(batch, dtype=np.int32)

            def e2e_step():
                N.check(N.lib().evdb_store_search_f64(
                    st._dev.handle, qn.ctypes.data_as(C.POINTER(C.c_double)), batch, DIM, K, N.COSINE,
                    slots.ctypes.data_as(C.POINTER(C.c_uint32)), dists.ctypes.data_as(C.POINTER(C.c_double)),
                    counts.ctypes.data_as(C.POINTER(C.c_int32))), "evdb_store_search_f64")
        else:
            def e2e_step():
                q2 = qh.to(dev, non_blocking=True)
                o = st.search(q2, K, "cosine")
                h_ids.copy_(o[0], non_blocking=True)
                h_d.copy_(o[1], non_blocking=True)
                torch.cuda.synchronize()
        for _ in range(min(warmup, 3)):
            e2e_step()
        barrier()
        lat = []
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            e2e_step()
            lat.append(time.perf_counter() - t1)
        barrier()
        e2e_s = time.perf_counter() - t0
        lat.sort()
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        return {"ms_total": ms, "ms_per_step": ms / steps, "qps": batch * steps / (ms / 1e3),
                "kernel_ms": (kms / nsamp) if nsamp else None, "kernel_samples": nsamp, "plan": plan,
                "launches": launches, "clocks": clocks, "flagged": flagged,
                "e2e_qps": batch * steps / e2e_s, "e2e_ms_per_step": e2e_s / steps * 1e3,
                "lat_p50_ms": lat[len(lat) // 2] * 1e3, "lat_p99_ms": lat[min(len(lat) - 1, int(len(lat) * 0.99))] * 1e3,
                "h2d": batch * DIM * 8, "d2h": batch * K * 16 + batch * 8}

    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full captures
    if os.path.exists(tp):
        traffic = json.load(open(tp))

    def roofline(r, batch):
        rows_local = st.hi - st.lo
        if r["kernel_ms"] is None:
            return None
        if r["plan"] == N.PLAN_GEMM and batch < 128:
            # a handful of queries: 2*B flop per operand byte is far below the machine balance, the
            # tcgen05 candidate pass is bound by reading its 2-byte operand column once
            byts = float(rows_local) * ((DIM + 63) // 64 * 64) * 2
            ach = byts / (r["kernel_ms"] * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "traffic": traffic.get("gemm_topk_kernel_b1") if world == 1 else None,
                    "algorithmic_bytes": byts, "peak_source": pk["src"] + " copy bandwidth",
                    "frac_of_8TBs_nominal": ach / 8000.0,
                    "kernel": "gemm_topk_kernel (fp16 operand column read once; candidates re-ranked in fp64 from the fp32 rows)",
                    "kernel_ms": r["kernel_ms"],
                    "fp32_row_bytes_equivalent_gbs": float(rows_local) * DIM * 4 / (r["kernel_ms"] * 1e-3) / 1e9}
        if r["plan"] == N.PLAN_GEMM:
            flops = 2.0 * rows_local * DIM * batch
            ach = flops / (r["kernel_ms"] * 1e-3) / 1e12
            return {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sustained"], "traffic": traffic.get("gemm_topk_kernel") if world == 1 else None,
                    "algorithmic_flops": flops, "peak_source": pk["src"] + " bf16 sustained",
                    "frac_of_burst": ach / pk["tf_burst"], "kernel": "gemm_topk_kernel (tcgen05 kind::f16, fp32 accumulate in TMEM)",
                    "kernel_ms": r["kernel_ms"]}
        byts = float(rows_local) * DIM * 4 * batch  # one corpus pass per query in the scan plan
        ach = byts / (r["kernel_ms"] * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": traffic.get("scan_float_kernel") if world == 1 else None,
                "algorithmic_bytes": byts, "peak_source": pk["src"] + " copy bandwidth",
                "frac_of_8TBs_nominal": ach / 8000.0, "kernel": "scan_float_kernel", "kernel_ms": r["kernel_ms"]}

    main = timed(args.batch, args.steps, args.warmup, sample_clocks=True)
    one = timed(BATCH_ONE, max(args.steps * 10, 50), max(args.warmup, 5), sample_clocks=False) \
        if args.batch != BATCH_ONE else None

    # secondary (N > 1): whole-store replicas answering disjoint query blocks (the
    # reference's own scale-out model) -- the throughput alternative to row sharding for a store
    # that fits one GPU.  Reported beside the row-sharded headline, never instead of it.
    replicas = None
    if world > 1:
        from erlvectordb_b200.sharded import ReplicaGroup
        rg = ReplicaGroup(dtype="f32", device=local, rank=rank, world=world)
        rg.fill_synthetic(synth.SEED_CORPUS, N_ROWS, DIM)
        qd = torch.from_numpy(synth.synth(synth.SEED_QUERY, 0, args.batch, DIM)).to(dev)
        for _ in range(args.warmup):
            rg.search(qd, K, "cosine")
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            rg.search(qd, K, "cosine")
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        replicas = {"value": args.batch * args.steps / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms / args.steps,
                    "note": f"{world} whole-store replicas, batch {args.batch} split into disjoint blocks, results all-gathered"}
        rg.close()

    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sample_rows = 100_000
        qps, t = cpu_reference_qps(threads, sample_rows, 3, 1)
        qps1, _ = cpu_reference_qps(1, sample_rows, 2, 1)   # one gen_server serialising one store
        cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "single_thread_value": qps1,
               "sample": f"{threads} threads x 1 query per step, 3 steps, first {sample_rows} rows of the "
                         f"1Mx768 corpus, QPS scaled by {sample_rows}/{N_ROWS}"}

    if rank == 0:
        out = {
            "metric": METRIC_NAME if args.batch == BATCH_MAIN else f"QPS (k=10, 1Mx768 cosine, batch {args.batch})",
            "value": main["qps"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if main["plan"] != N.PLAN_GEMM else "f16 (tcgen05 candidates, fp32 accumulate) + f64 exact re-rank",
            "data": "synthetic",
            "config": {"workload": "1Mx768 fp32 cosine k=10 (BASELINE.json configs[1])", "rows": N_ROWS,
                       "dim": DIM, "k": K, "batch": args.batch, "plan": {1: "scan", 2: "gemm", 3: "exact"}.get(main["plan"]),
                       "sharding": f"rows/{world}" if world > 1 else "none",
                       "cache": "inputs larger than L2 (3.07 GB fp32 corpus, 126 MB L2)",
                       "result_check": "candidates re-ranked in exact fp64 on device; ids+distances bit-equal to the oracle in tests/"},
            "clocks": main["clocks"],
            "e2e": {"value": main["e2e_qps"], "unit": "queries/s", "h2d_bytes_per_step": main["h2d"],
                    "d2h_bytes_per_step": main["d2h"], "ms_per_step": main["e2e_ms_per_step"]},
            "gpu_launches": main["launches"],
            "roofline": roofline(main, args.batch),
            "cpu_baseline": cpu,
            "escalated_queries": main["flagged"],
        }
        if replicas is not None:
            out["replica_groups"] = replicas
        if one is not None:
            out["batch1"] = {"value": one["qps"], "unit": "queries/s", "ms_per_step": one["ms_per_step"],
                             "latency_ms": {"p50": one["lat_p50_ms"], "p99": one["lat_p99_ms"],
                                            "how": "wall clock around the host-buffer call"},
                             "e2e": {"value": one["e2e_qps"], "unit": "queries/s",
                                     "h2d_bytes_per_step": one["h2d"], "d2h_bytes_per_step": one["d2h"]},
                             "roofline": roofline(one, BATCH_ONE), "gpu_launches": one["launches"],
                             "steps": max(args.steps * 10, 50)}
        print(json.dumps(out))
    st.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_MAIN)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
