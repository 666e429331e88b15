#!/usr/bin/env python
"""bench.py -- QPS of the search hot path on the BASELINE.json workload.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): 1,000,000 x 768 fp32 synthetic corpus, cosine, k = 10,
query batch 1024 (the headline `value`, tensor-bound) and batch 1 (reported under "batch1",
HBM-bound).  A step is one search call over one batch; consecutive steps use DIFFERENT query
batches (a pool of 4).  N > 1 row-shards the SAME corpus over the ranks (strong scaling), one
process per GPU: each GPU searches its shard, candidates meet over peer memory (NVLink).

`value`   : device-timed (CUDA events), queries already resident in HBM, max over ranks.
`e2e`     : the same metric through the reference-facing C-ABI call evdb_store_search_f64 with
            PAGEABLE host buffers (what a NIF binary is): host->device query copy and
            device->host result copy inside the timed region.  N = 1: the rank's own store; N > 1:
            ONE handle over all N devices (evdb_opts.n_shards, mstore.cu) driven by rank 0's
            process -- the way one BEAM process would drive the box.
`roofline`: dominant kernel (scan: HBM bytes; tcgen05 GEMM: flops) timed with CUDA events on its
            launching stream inside the timed region, against MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (a port of the reference's Erlang arithmetic: full fp64 scan with
            the query norm recomputed per row + full sort) on a bounded sample, rank 0, N = 1.
`config.result_check`: after the timed region the results of every batch size are compared bit for
            bit with a single-device store's on rank 0 and their digest is printed -- the same digest
            must appear at N = 1, 2, 4, 8.
`configs` : the other BASELINE.json configurations (cfg1, cfg3, cfg4_weak, cfg5) as secondary lines,
            each with its roofline and result digest (skipped with --no-extra).
The corpus (3.07 GB; 1.5 GB fp16 operand column) is far larger than the 126 MB L2, so consecutive
steps cannot be served from cache ("inputs larger than L2").
"""
from __future__ import annotations

import argparse
import hashlib
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
POOL = 4   # distinct query batches rotated through the timed steps
METRIC_NAME = "QPS (k=10, 1Mx768 cosine, batch 1024)"
BPR = {"f32": lambda d: 4 * d, "bf16": lambda d: 2 * d, "u8": lambda d: d + 8, "u4": lambda d: (d + 1) // 2 + 8}


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
def digest(ids, dists) -> str:
    import numpy as np
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(ids).astype(np.uint32).tobytes())
    h.update(np.ascontiguousarray(dists, dtype=np.float64).tobytes())
    return h.hexdigest()[:16]


def run_ours(args):
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist
    from erlvectordb_b200 import _native as N
    from erlvectordb_b200 import synth
    from erlvectordb_b200.device_store import DeviceStore
    from erlvectordb_b200.sharded import ShardedStore

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; erlvectordb_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # a HOST barrier for the sections where only rank 0 drives the GPUs (an NCCL barrier would park a
        # spinning kernel on every other GPU for the whole measurement)
        cpu_group = dist.new_group(backend="gloo")
    rc = N.lib().evdb_init(None, 0)
    if rc != 0:
        raise SystemExit(f"evdb_init: {N.lib().evdb_strerror(rc).decode()}")
    pk = peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full captures
    if os.path.exists(tp):
        traffic = json.load(open(tp))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def c_abi_search(store: DeviceStore, qn, k, metric, slots, dists, counts):
        N.check(N.lib().evdb_store_search_f64(
            store.handle, qn.ctypes.data_as(C.POINTER(C.c_double)), qn.shape[0], qn.shape[1], k, N.METRICS[metric],
            slots.ctypes.data_as(C.POINTER(C.c_uint32)), dists.ctypes.data_as(C.POINTER(C.c_double)),
            counts.ctypes.data_as(C.POINTER(C.c_int32))), "evdb_store_search_f64")

    def roofline(plan, kernel_ms, batch, rows_local, d, dtype, tkey=None):
        if kernel_ms is None or kernel_ms <= 0:
            return None
        tr = traffic.get(tkey) if (tkey and world == 1) else None
        if plan == N.PLAN_GEMM and dtype == "u8":
            # quantization_8bit batches: tcgen05 kind::i8 over the codes x two query digit planes (gemm_i8.cu).  One logical
            # dot product = 2*d integer ops per (query, row); the kernel issues two MMAs (high and low digit) for it.  There is
            # no measured int8 peak in MEASURED_PEAKS.json: the bf16 figure is the denominator, and says so.
            ops = 2.0 * rows_local * d * batch
            ach = ops / (kernel_ms * 1e-3) / 1e12
            return {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sustained"], "traffic": tr, "algorithmic_flops": ops,
                    "peak_source": pk["src"] + " bf16 sustained (no measured int8 peak; the kernel runs 2 digit-plane MMAs per logical dot)",
                    "frac_of_burst": ach / pk["tf_burst"],
                    "kernel": "gemm_i8_topk_kernel (tcgen05 kind::i8, int32 accumulate in TMEM, exact digit-plane sums)",
                    "kernel_ms": kernel_ms,
                    "code_bytes_equivalent_gbs": float(rows_local) * BPR[dtype](d) * batch / (kernel_ms * 1e-3) / 1e9}
        if plan == N.PLAN_GEMM and batch < 128:
            # a handful of queries: 2*B flop per operand byte is far below the machine balance, the
            # tcgen05 candidate pass is bound by reading its 2-byte operand column once
            byts = float(rows_local) * ((d + 63) // 64 * 64) * 2
            ach = byts / (kernel_ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": tr, "algorithmic_bytes": byts, "peak_source": pk["src"] + " copy bandwidth",
                    "frac_of_8TBs_nominal": ach / 8000.0,
                    "kernel": "gemm_topk_kernel (fp16 operand column read once; candidates re-ranked in fp64 from the fp32 rows)",
                    "kernel_ms": kernel_ms,
                    "fp32_row_bytes_equivalent_gbs": float(rows_local) * d * 4 / (kernel_ms * 1e-3) / 1e9}
        if plan == N.PLAN_GEMM:
            flops = 2.0 * rows_local * d * batch
            ach = flops / (kernel_ms * 1e-3) / 1e12
            return {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sustained"], "traffic": tr, "algorithmic_flops": flops,
                    "peak_source": pk["src"] + " bf16 sustained", "frac_of_burst": ach / pk["tf_burst"],
                    "kernel": "gemm_topk_kernel (tcgen05 kind::f16, fp32 accumulate in TMEM)", "kernel_ms": kernel_ms}
        byts = float(rows_local) * BPR[dtype](d)   # one corpus pass serves the whole launch
        ach = byts / (kernel_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": tr, "algorithmic_bytes": byts, "peak_source": pk["src"] + " copy bandwidth",
                "frac_of_8TBs_nominal": ach / 8000.0, "kernel": "scan kernel of the plan", "kernel_ms": kernel_ms}

    def device_timed(st: ShardedStore, n_rows, d, k, metric, batch, steps, warmup, sample_clocks=False, phases=False,
                     preload=False):
        """K steps of st.search over rotating resident query batches, CUDA events, max over ranks."""
        qd = [torch.from_numpy(synth.synth(synth.SEED_QUERY, i * batch, batch, d)).to(dev) for i in range(POOL)]
        for i in range(warmup):
            st.search(qd[i % POOL], k, metric, escalate=False)
        barrier()
        sampler = ClockSampler(local)
        if preload and not sample_clocks:   # the same 0.6 s of load, unsampled: both layouts are timed in the sustained state
            t_probe = time.perf_counter()
            i = 0
            while time.perf_counter() - t_probe < 0.6:
                for _ in range(8):
                    st.search(qd[i % POOL], k, metric, escalate=False)
                    i += 1
                torch.cuda.synchronize()
        if sample_clocks:
            # nvidia-smi samples every 200 ms and the timed region may last only tens of ms: the
            # sampler also covers ~0.6 s of the SAME search loop run (untimed) right before it, so
            # that the clocks/throttle record describes this load, and the timed steps start from
            # the steady (power-capped) state rather than from a cold burst
            sampler.start()
            t_probe = time.perf_counter()
            i = 0
            while time.perf_counter() - t_probe < 0.6:
                for _ in range(8):
                    st.search(qd[i % POOL], k, metric, escalate=False)
                    i += 1
                torch.cuda.synchronize()
        # the dominant kernel is bracketed by the library's event pairs INSIDE the timed steps (roofline contract); the
        # two event records per search sit between kernels of the chain, so the device-timed loop runs without the
        # programmatic dependent launch of those two links (the host-buffer path below has it)
        st._dev.profile(True)
        if phases and hasattr(st, "phase_events"):
            st.phase_events = []
        l0 = st._dev.stats()["kernel_launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            # enqueue only: the per-call flag read-back of the escalating API would put a host round trip
            # between steps; every distinct batch goes through that API in result_check below
            st.search(qd[i % POOL], k, metric, escalate=False)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if sample_clocks else None
        launches_timed = st._dev.stats()["kernel_launches"] - l0
        nsamp, kms = st._dev.profile_read()
        st._dev.profile(False)
        ph = None
        if phases and getattr(st, "phase_events", None):
            nph = len(st.phase_events[0]) - 1
            acc = [0.0] * nph
            for ev in st.phase_events:
                for j in range(nph):
                    acc[j] += ev[j].elapsed_time(ev[j + 1])
            n = len(st.phase_events)
            names = (["phase1_prep_seed_gemm_window_push_us", "phase2_merge_rerank_push_us", "phase3_final_us"] if nph == 3
                     else ["local_search_prep_seed_gemm_select_rerank_us", "push_wait_merge_us"])
            ph = {nm: max_over_ranks(acc[j] / n * 1e3) for j, nm in enumerate(names)}
            ph["gemm_kernel_us"] = max_over_ranks((kms / nsamp) * 1e3 if nsamp else 0.0)
            ph["scheme"] = ("two-phase: approximate windows travel, owners re-rank" if nph == 3
                            else "one exchange: every rank re-ranks its local window, finished results travel")
            ph["how"] = "CUDA events between the phases on the real GPUs, mean over the timed steps, max over ranks"
        if hasattr(st, "phase_events"):
            st.phase_events = None
        stt = st._dev.stats()
        launches = launches_timed + (2 * steps if world > 1 and hasattr(st, "exchange") else 0)  # + the exchange push and merge kernels
        # every distinct batch through the escalating API: nothing unproven may be left
        esc0 = st.n_escalations
        for i in range(POOL):
            out = st.search(qd[i], k, metric)
            assert int(out[3].sum().item()) == 0
        return {"ms_per_step": ms / steps, "qps": batch * steps / (ms / 1e3),
                "kernel_ms": (kms / nsamp) if nsamp else None, "plan": stt["last_plan"], "launches": launches,
                "clocks": clocks, "phases": ph, "escalated": st.n_escalations - esc0, "qd": qd,
                "exchange": ("none" if world == 1 else "NCCL all_gather_into_tensor of the replicas' result blocks"
                             if not hasattr(st, "exchange") else
                             "peer-memory mailboxes (CUDA IPC, NVLink stores, on-device flag wait)" if st.exchange == "p2p"
                             else "NCCL all_gather_into_tensor + merge kernel")}

    def e2e_timed(store: DeviceStore, d, k, metric, batch, steps, warmup):
        """The C-ABI host call with pageable numpy buffers, wall clock, rank 0's process."""
        qn = [np.array(synth.synth(synth.SEED_QUERY, i * batch, batch, d)) for i in range(POOL)]   # pageable
        slots = np.empty((batch, k), dtype=np.uint32)
        dists = np.empty((batch, k), dtype=np.float64)
        counts = np.empty(batch, dtype=np.int32)
        for i in range(max(3, min(warmup, 5))):
            c_abi_search(store, qn[i % POOL], k, metric, slots, dists, counts)
        lat, h2d, dv, d2h = [], [], [], []
        t0 = time.perf_counter()
        for i in range(steps):
            t1 = time.perf_counter()
            c_abi_search(store, qn[i % POOL], k, metric, slots, dists, counts)
            lat.append(time.perf_counter() - t1)
        total = time.perf_counter() - t0
        for i in range(min(steps, 8)):   # the library's own event split of a call (not in the timed loop)
            c_abi_search(store, qn[i % POOL], k, metric, slots, dists, counts)
            s = store.stats()
            h2d.append(s["last_h2d_ms"]); dv.append(s["last_device_ms"]); d2h.append(s["last_d2h_ms"])
        lat.sort()
        return {"qps": batch * steps / total, "ms_per_step": total / steps * 1e3,
                "p50_ms": lat[len(lat) // 2] * 1e3, "p99_ms": lat[min(len(lat) - 1, int(len(lat) * 0.99))] * 1e3,
                "h2d": batch * d * 8, "d2h": batch * k * 12 + batch * 4,
                "breakdown_ms": {"h2d": statistics.mean(h2d), "device": statistics.mean(dv), "d2h": statistics.mean(d2h),
                                 "how": "CUDA events inside the library around the query copy, the device work and the result copy"
                                        + ("; this call is replayed as ONE CUDA graph whose nodes include both copies, so the "
                                           "whole graph is reported under 'device'" if statistics.mean(h2d) == 0.0 else "")}}

    def check_results(st: ShardedStore, single: DeviceStore | None, multi: DeviceStore | None, d, k, metric, batch):
        """Sharded result of pool batch 0 (escalating API) == single-device store == one-handle multi-device store."""
        qh = np.array(synth.synth(synth.SEED_QUERY, 0, batch, d))
        out = st.search(torch.from_numpy(qh).to(dev), k, metric)
        ids, dd = out[0].cpu().numpy(), out[1].cpu().numpy()
        mine = digest(ids, dd)
        digs = [mine]
        if world > 1:
            digs = [None] * world
            dist.all_gather_object(digs, mine)
        res = {"digest": mine, "all_ranks_equal": len(set(digs)) == 1}
        if rank == 0 and single is not None:
            s_ids, s_d, _ = single.search(qh, k, metric)
            res["single_gpu_digest"] = digest(s_ids, s_d)
            res["identical_to_single_gpu"] = bool(np.array_equal(ids.astype(np.uint32), s_ids) and np.array_equal(dd, s_d))
        if rank == 0 and multi is not None:
            m_ids, m_d, _ = multi.search(qh, k, metric)
            res["one_handle_multi_device_digest"] = digest(m_ids, m_d)
        return res

    # ================================ headline: configs[1] ================================
    # N > 1: the corpus is row-sharded (the headline).  Whole-store replicas answering disjoint query blocks (the
    # reference's own scale-out model) are measured beside it in the SAME sustained state and reported
    # under `layouts` -- never instead of it.
    from erlvectordb_b200.sharded import ReplicaGroup
    layout = "rows"
    st = ShardedStore(dtype="f32", device=local, rank=rank, world=world)
    st.fill_synthetic(synth.SEED_CORPUS, N_ROWS, DIM)
    rows_local = st.hi - st.lo
    main_rows = device_timed(st, N_ROWS, DIM, K, "cosine", args.batch, args.steps, args.warmup, sample_clocks=layout == "rows",
                             phases=world > 1, preload=True)
    main_rep = None
    if world > 1:
        rg = ReplicaGroup(dtype="f32", device=local, rank=rank, world=world)
        rg.fill_synthetic(synth.SEED_CORPUS, N_ROWS, DIM)
        main_rep = device_timed(rg, N_ROWS, DIM, K, "cosine", args.batch, args.steps, args.warmup, sample_clocks=layout == "replicas",
                                preload=True)
        r = rg.search(main_rep["qd"][0], K, "cosine")
        main_rep["digest"] = digest(r[0].cpu().numpy(), r[1].cpu().numpy())
        rg.close()
        del rg
    main = main_rep if layout == "replicas" else main_rows
    one = None
    if args.batch != BATCH_ONE:
        one = device_timed(st, N_ROWS, DIM, K, "cosine", BATCH_ONE, max(args.steps * 10, 50), max(args.warmup, 5))

    # the stores the reference-facing call is measured on (and the results are checked against)
    single = multi = None
    if rank == 0:
        if world == 1:
            single = st._dev
        else:
            single = DeviceStore(dtype="f32", device=local)
            single.fill_synthetic(synth.SEED_CORPUS, N_ROWS, DIM)
            multi = DeviceStore(dtype="f32", devices=list(range(world)))
            multi.fill_synthetic(synth.SEED_CORPUS, N_ROWS, DIM)
    host_barrier()
    e2e_main = e2e_one = None
    if rank == 0:
        host_store = single if world == 1 else multi
        e2e_main = e2e_timed(host_store, DIM, K, "cosine", args.batch, args.steps, args.warmup)
        if one is not None:
            e2e_one = e2e_timed(host_store, DIM, K, "cosine", BATCH_ONE, max(args.steps * 10, 50), 5)
    host_barrier()
    chk_main = check_results(st, single, multi, DIM, K, "cosine", args.batch)
    chk_one = check_results(st, single, multi, DIM, K, "cosine", BATCH_ONE) if one is not None else None

    if multi is not None:
        multi.close()
    if single is not None and world > 1:
        single.close()
    st.close()
    del st
    torch.cuda.empty_cache()

    # ================================ the other BASELINE configurations ================================
    def extra(name, n_total, d, dtype, metric, k, batch, steps, note, weak=False):
        s2 = ShardedStore(dtype=dtype, device=local, rank=rank, world=world)
        s2.fill_synthetic(synth.SEED_CORPUS, n_total, d)
        r = device_timed(s2, n_total, d, k, metric, batch, steps, 3)
        ref = None
        if rank == 0:
            ref = s2._dev if world == 1 else DeviceStore(dtype=dtype, device=local)
            if world > 1:
                ref.fill_synthetic(synth.SEED_CORPUS, n_total, d)
        chk = check_results(s2, ref, None, d, k, metric, batch)
        e2e = None
        if rank == 0 and batch == 1:
            e2e = e2e_timed(ref, d, k, metric, 1, max(steps, 50), 5)
        e2e_batch = None
        if rank == 0 and world == 1 and batch > 1 and dtype == "u8":   # the i8 plan through the C ABI with host buffers
            e2e_batch = e2e_timed(ref, d, k, metric, batch, max(steps, 5), 3)
        rl = roofline(r["plan"], r["kernel_ms"], batch, s2.hi - s2.lo, d, dtype)
        rec = {"workload": name, "rows": n_total, "dim": d, "dtype": dtype, "metric": metric, "k": k, "batch": batch,
               "scaling": "weak" if weak else "strong", "value": r["qps"], "unit": "queries/s", "ms_per_step": r["ms_per_step"],
               "plan": {1: "scan", 2: "gemm", 3: "exact"}.get(r["plan"]), "roofline": rl, "result_check": chk,
               "escalated_queries": r["escalated"], "note": note}
        if e2e is not None:
            rec["latency_ms_single_gpu_c_abi"] = {"p50": e2e["p50_ms"], "p99": e2e["p99_ms"]}
        if e2e_batch is not None:
            rec["e2e"] = {"value": e2e_batch["qps"], "unit": "queries/s", "h2d_bytes_per_step": e2e_batch["h2d"],
                          "d2h_bytes_per_step": e2e_batch["d2h"], "ms_per_step": e2e_batch["ms_per_step"],
                          "breakdown_ms": e2e_batch["breakdown_ms"], "through": "evdb_store_search_f64, pageable host buffers"}
        if rank == 0 and world > 1:
            ref.close()
        s2.close()
        del s2
        torch.cuda.empty_cache()
        return rec

    configs = None
    if not args.no_extra:
        configs = {
            "cfg1": extra("configs[0]: 10k x128 cosine k=10, batch 1 (the reference's own CPU-runnable case)",
                          10_000, 128, "f32", "cosine", 10, 1, 200, "L2-resident, latency-bound: no roofline claim"),
            "cfg3": extra("configs[2]: 10M x128 euclidean k=100, batch 4096 (tcgen05 GEMM + top-k epilogue)",
                          10_000_000, 128, "f32", "euclidean", 100, 4096, max(3, min(args.steps, 6)),
                          "row norms ride in the K dimension; fp64 re-rank of the squared-distance candidates"),
            "cfg4_weak": extra(f"configs[3]: {12.5 * world:g}M x96 quantization_8bit over {world} GPU(s), 12.5M rows per GPU, batch 1",
                               12_500_000 * world, 96, "u8", "cosine", 10, 1, 100,
                               "int8 dp4a scan on TMA-staged tiles + peer-memory exchange + merge; weak scaling: north_star's 100M point is N = 8",
                               weak=True),
            "cfg4_batch": extra(f"configs[3] row shape, query batch: {12.5 * world:g}M x96 quantization_8bit over {world} GPU(s), batch 1024",
                                12_500_000 * world, 96, "u8", "cosine", 10, 1024, max(3, min(args.steps, 5)),
                                "tcgen05 kind::i8 over the stored codes x the query's two 8-bit digit planes (exact integer sums), "
                                "fused top-k epilogue, fp64 re-rank from the codes; the dp4a scan serves one query per pass",
                                weak=True),
            "cfg5_manhattan": extra("configs[4]: 1M x1536 manhattan k=10, batch 1", 1_000_000, 1536, "f32", "manhattan", 10, 1, 50,
                                    "bandwidth-bound fp32 scan"),
            "cfg5_u4": extra("configs[4]: 1M x1536 quantization_4bit cosine k=10, batch 1", 1_000_000, 1536, "u4", "cosine", 10, 1, 100,
                             "bandwidth-bound packed-nibble scan"),
        }

    # ---- f-3: inserts and deletes at rate through the C ABI (rank 0, one device) ----
    ingest = None
    if rank == 0 and not args.no_extra:
        ing = DeviceStore(dtype="f32", device=local, capacity_hint=70_000)
        base = np.array(synth.synth(synth.SEED_CORPUS, 0, 4096, DIM))             # pageable fp64 rows, as a NIF sees them
        ptrs = [base[i].ctypes.data_as(C.POINTER(C.c_double)) for i in range(4096)]   # (the ctypes casts are not the product)
        up, de = N.lib().evdb_store_upsert_f64, N.lib().evdb_store_delete
        for i in range(64):                                                        # warm-up: ring, workspaces
            up(ing.handle, i, ptrs[i], DIM)
        ing.append(base); ing.flush()
        c0 = ing.stats()["count"]
        n_up = 20_000
        t0 = time.perf_counter()
        for i in range(n_up):
            up(ing.handle, c0 + i, ptrs[i & 4095], DIM)
        ing.flush()
        t_up = time.perf_counter() - t0
        t0 = time.perf_counter()
        for i in range(10):
            ing.append(base)
        ing.flush()
        t_app = time.perf_counter() - t0
        n_del = 20_000
        cnt = ing.stats()["count"]
        moved = C.c_int64()
        t0 = time.perf_counter()
        for i in range(n_del):
            de(ing.handle, (i * 7919) % (cnt - i), C.byref(moved))
        ing.flush()
        t_del = time.perf_counter() - t0
        ingest = {"upserts_per_s": n_up / t_up, "append_rows_per_s": 40960 / t_app, "deletes_per_s": n_del / t_del,
                  "rows": "768 x fp64 from pageable host memory, fp32 store with the fp16 operand column",
                  "how": "evdb_store_upsert_f64 one row per call (enqueue only: pinned staging ring, narrow + finalize kernels), "
                         "evdb_store_append_f64 4096 rows per call, evdb_store_delete (one swap kernel); flushed at the end of each leg"}
        ing.close()

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
                       "sharding": ("none" if world == 1 else f"rows/{world}" if layout == "rows"
                                    else f"{world} whole-store replicas x {args.batch // world} queries each"),
                       "exchange": main["exchange"],
                       "queries": f"{POOL} distinct batches rotated through the steps",
                       "cache": "inputs larger than L2 (3.07 GB fp32 corpus, 126 MB L2)",
                       "result_check": chk_main},
            "clocks": main["clocks"],
            "e2e": {"value": e2e_main["qps"], "unit": "queries/s", "h2d_bytes_per_step": e2e_main["h2d"],
                    "d2h_bytes_per_step": e2e_main["d2h"], "ms_per_step": e2e_main["ms_per_step"],
                    "breakdown_ms": e2e_main["breakdown_ms"],
                    "through": "evdb_store_search_f64, pageable host buffers" +
                               ("" if world == 1 else f", one handle over {world} devices (evdb_opts.n_shards) in rank 0's process")},
            "gpu_launches": main["launches"],
            "roofline": roofline(main["plan"], main["kernel_ms"], args.batch if layout == "rows" else args.batch // world,
                                 rows_local if layout == "rows" else N_ROWS, DIM, "f32", "gemm_topk_kernel"),
            "cpu_baseline": cpu,
            "escalated_queries": main["escalated"],
        }
        if main["phases"]:
            out["phases"] = main["phases"]
        if main_rep is not None:
            lay = lambda r, note: {"value": r["qps"], "unit": "queries/s", "ms_per_step": r["ms_per_step"], "note": note}
            out["layouts"] = {"chosen": layout,
                              "rows": lay(main_rows, f"the corpus row-sharded over {world} GPUs, every GPU sees every query, one exchange + merge"),
                              "replicas": lay(main_rep, f"{world} whole-store replicas, batch {args.batch} split into disjoint blocks, results all-gathered")}
            out["layouts"]["replicas"]["digest"] = main_rep["digest"]
            out["layouts"]["rows"]["phases"] = main_rows["phases"]
            out["replica_groups"] = out["layouts"]["replicas"]
        if one is not None:
            out["batch1"] = {"value": one["qps"], "unit": "queries/s", "ms_per_step": one["ms_per_step"],
                             "latency_ms": {"p50": e2e_one["p50_ms"], "p99": e2e_one["p99_ms"],
                                            "how": "wall clock around the host-buffer C-ABI call"},
                             "e2e": {"value": e2e_one["qps"], "unit": "queries/s", "h2d_bytes_per_step": e2e_one["h2d"],
                                     "d2h_bytes_per_step": e2e_one["d2h"], "breakdown_ms": e2e_one["breakdown_ms"]},
                             "roofline": roofline(one["plan"], one["kernel_ms"], BATCH_ONE, rows_local, DIM, "f32",
                                                  "gemm_topk_kernel_b1" if one["plan"] == N.PLAN_GEMM else "scan_float_kernel"),
                             "gpu_launches": one["launches"], "steps": max(args.steps * 10, 50), "result_check": chk_one}
        if configs is not None:
            out["configs"] = configs
        if ingest is not None:
            out["ingest"] = ingest
        print(json.dumps(out))
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
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary BASELINE configurations")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
