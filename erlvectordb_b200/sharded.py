"""Row-sharded store: one process per GPU, candidates merged after an NCCL allgather.

The reference has no sharded search (cluster_manager replicates whole stores,
reference src/cluster_manager.erl:148-171); rows are independent and top-k is an
associative merge, so the corpus splits by contiguous row blocks:

    rank r owns global rows [lo_r, hi_r);  every rank sees every query
    local:   evdb_store_search_dev  -> (global id u64, exact fp64 distance) x k
    exchange: ONE torch.distributed all_gather_into_tensor over NCCL/NVLink of the packed
             per-rank result blob (B*k*16 B + B*8 B per rank)
    merge:   evdb_merge_topk_packed_dev on every rank -> identical global top-k everywhere

Because each shard's distances are the exact fp64 values and ids are global rows,
the merged result is bit-identical to the single-GPU result.

torch is plumbing here (device buffers, streams, the process group); the scan,
re-rank and merge are the library's CUDA kernels.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import _native as N
from .device_store import (DeviceStore, Exchange, gemm_window, merge_topk_packed_dev, sharded_phase1,
                           sharded_phase2, sharded_phase3)


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row blocks of ceil(n/world) rows; the last shards may be short or empty."""
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    hi = min(n_total, lo + per)
    return lo, hi


def blob_words(B: int, k: int) -> int:
    """64-bit words of one rank's packed result: [B*k ids][B*k dists][B counts i32][B flags i32]."""
    return 2 * B * k + B


def blob_views(blob: torch.Tensor, B: int, k: int):
    """(ids (B,k) int64, dists (B,k) float64, counts (B,) int32, flags (B,) int32) views of a blob."""
    nk = B * k
    ids = blob[:nk].view(B, k)
    dists = blob[nk:2 * nk].view(torch.float64).view(B, k)
    cf = blob[2 * nk:].view(torch.int32)
    return ids, dists, cf[:B], cf[B:]


def gather_blobs(blob: torch.Tensor, world: int, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """ONE collective per search: all_gather every rank's packed blob -> (world, words)."""
    if out is None:
        out = torch.empty((world, blob.numel()), dtype=blob.dtype, device=blob.device)
    if world == 1:
        out[0].copy_(blob)
    elif blob.is_cuda:  # NCCL over NVLink
        dist.all_gather_into_tensor(out.view(-1), blob, group=group)
    else:               # gloo (CPU tests of the host logic)
        parts = [torch.empty_like(blob) for _ in range(world)]
        dist.all_gather(parts, blob, group=group)
        for g, p_ in enumerate(parts):
            out[g].copy_(p_)
    return out


def _stream_handle(dev) -> int:
    """cudaStream_t of torch's current stream for the C ABI.  The ABI reads NULL as "the store's
    own stream", so torch's default stream (handle 0) is passed as cudaStreamLegacy (0x1)."""
    h = torch.cuda.current_stream(dev).cuda_stream
    return h if h else 1


class ShardedStore:
    """One shard of a row-sharded store (call from every rank of the process group)."""

    def __init__(self, dtype="f32", device=0, rank=None, world=None, group=None,
                 local_search=None, merge=None, exchange="p2p"):
        """exchange: "p2p" = peer-memory pushes + on-device flag wait (evdb_exchange_*, the default on
        GPUs; falls back to "nccl" if the peers cannot be mapped), "nccl" = one all_gather_into_tensor
        followed by the merge kernel."""
        self.exchange = exchange
        self._xchg = {}
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.device = device
        self.dtype = dtype
        self._local_search = local_search   # None -> the CUDA library
        self._merge = merge                 # None -> the CUDA merge kernel
        self._dev = None if local_search else DeviceStore(dtype=dtype, device=device)
        self._bufs = {}
        self.lo = self.hi = 0
        self.n_total = 0
        self.dim = 0
        self.n_escalations = 0
        self.phase_events = None   # a list: the two-phase search records 4 CUDA events per call into it

    # -- ingest --------------------------------------------------------------------
    def fill_synthetic(self, seed: int, n_total: int, d: int):
        self.n_total, self.dim = n_total, d
        self.lo, self.hi = shard_bounds(n_total, self.world, self.rank)
        if self._dev is not None and self.hi > self.lo:
            self._dev.fill_synthetic(seed, self.hi - self.lo, d, row0=self.lo)

    def bulk_load_shard(self, rows, n_total: int):
        """`rows`: this rank's block [lo, hi) of the global row-major corpus."""
        self.n_total, self.dim = n_total, rows.shape[1]
        self.lo, self.hi = shard_bounds(n_total, self.world, self.rank)
        assert rows.shape[0] == self.hi - self.lo
        if self._dev is not None and self.hi > self.lo:
            self._dev.bulk_load(rows)

    # -- search ----------------------------------------------------------------------
    def _buffers(self, B: int, k: int, dev):
        key = (B, k, dev)
        if key not in self._bufs:  # reused across searches: no allocator / fill kernels per call
            w = blob_words(B, k)
            local = torch.zeros((w,), dtype=torch.int64, device=dev)
            gathered = torch.zeros((self.world, w), dtype=torch.int64, device=dev)
            merged = torch.zeros((w,), dtype=torch.int64, device=dev)
            # views and raw pointers are made once: a lone query on a small store is host-bound
            lv, mv = blob_views(local, B, k), blob_views(merged, B, k)
            self._bufs[key] = (local, gathered, merged, lv, mv, tuple(t.data_ptr() for t in lv))
        return self._bufs[key]

    def _cuda_local_search(self, q: torch.Tensor, k: int, metric: str, ptrs, plan="auto", kp_min=0):
        """The library writes this shard's result straight into the packed blob (ptrs: its four views)."""
        B, d = q.shape
        if self.hi > self.lo:
            if plan == "auto" and kp_min == 0:
                self._dev.search_dev(q.data_ptr(), B, d, k, metric, self.lo, ptrs[0], ptrs[1], ptrs[2], ptrs[3],
                                     _stream_handle(q.device))
            else:
                self._dev.search_dev_ex(q.data_ptr(), B, d, k, metric, ptrs[0], ptrs[1], ptrs[2], ptrs[3],
                                        _stream_handle(q.device), plan=plan, kp_min=kp_min, slot_base=self.lo)
        # an empty shard keeps the zero counts the blob was created with

    def _p2p(self, B: int, k: int, dev, words: int | None = None, tag: str = "blob"):
        """An exchange object for (B, k), created and connected on first use (a collective call)."""
        key = (B, k, tag)
        if key not in self._xchg:
            x = None
            try:
                x = Exchange(dev.index or 0, self.rank, self.world, blob_words(B, k) if words is None else words)
                mine = torch.tensor(list(bytes(x.handle_bytes)), dtype=torch.uint8, device=dev)
                allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(allh.view(-1), mine, group=self.group)
                x.connect(bytes(allh.cpu().numpy().tobytes()))
                ok = torch.ones((1,), dtype=torch.int32, device=dev)
            except Exception:  # peers not mappable (no P2P / IPC): everyone must agree to fall back
                ok = torch.zeros((1,), dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                if x is not None:
                    x.close()
                x = None
                if self.exchange != "nccl" and self.rank == 0:
                    import sys
                    print("[erlvectordb_b200] peer mailboxes could not be mapped (no CUDA IPC / P2P between the ranks): "
                          "the sharded search exchanges through NCCL all_gather instead", file=sys.stderr)
                self.exchange = "nccl"
            self._xchg[key] = x
        return self._xchg[key]

    def _cuda_merge(self, gathered: torch.Tensor, B: int, k: int, merged: torch.Tensor):
        dev = gathered.device
        merge_topk_packed_dev(dev.index or 0, gathered.data_ptr(), self.world, B, k, merged.data_ptr(),
                              _stream_handle(dev))

    def _two_phase_ok(self, B: int, k: int, metric: str) -> bool:
        """Decided from GLOBAL facts only, so that every rank takes the same path: the tcgen05 plan
        must apply on every shard (F32, cosine/euclidean, a real batch, a window <= 128 keys, every
        shard at least one corpus tile)."""
        if self._merge is not None or self._local_search is not None or self.exchange != "p2p" or self.world < 2:
            return False
        # Since the exact fold runs on a group of warps per query (select.cu mw_fold) re-ranking a whole local
        # window costs a rank little more than re-ranking its eighth of the global one, and ONE exchange
        # (finished per-shard results + merge) beats the two of the window/owner scheme: N = 4, batch 1024:
        # 0.478 against 0.495 ms.  The two-phase search stays available (EVDB_SHARD_TWO_PHASE=1) and tested.
        if os.environ.get("EVDB_SHARD_TWO_PHASE", "0") != "1":
            return False
        per = (self.n_total + self.world - 1) // self.world
        smallest = self.n_total - per * (self.world - 1)
        return (self.dtype == "f32" and metric in ("cosine", "euclidean") and 16 <= B <= 8192 and
                gemm_window(k, self.n_total) <= 128 and self.world * gemm_window(k, self.n_total) <= 2048 and
                self.world <= 32 and      # the phase kernels map one lane per rank
                smallest >= 256 and self.n_total < 0xFFFFFFF0)

    def _search_two_phase(self, q: torch.Tensor, k: int, metric: str):
        """Query batches on the tcgen05 plan: approximate windows travel, owners re-rank (select.cu)."""
        B, d = q.shape
        kp = gemm_window(k, self.n_total)
        xw = self._p2p(B, k, q.device, words=B * kp + B, tag="win")
        xe = self._p2p(B, k, q.device, words=B * kp, tag="exact")
        if xw is None or xe is None:
            return None
        merged, mv = self._buffers(B, k, q.device)[2], self._buffers(B, k, q.device)[4]
        stream = _stream_handle(q.device)
        ev = None
        if self.phase_events is not None:   # measurement aid (bench.py): CUDA events between the phases
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        rc = sharded_phase1(self._dev, xw, q.data_ptr(), B, d, k, metric, self.lo, self.n_total, stream)
        if rc != N.OK:
            raise N.EvdbError(rc, "sharded_phase1")   # the plan was agreed on from global facts: must not fail alone
        if ev:
            ev[1].record()
        sharded_phase2(self._dev, xw, xe, q.data_ptr(), B, k, metric, self.n_total, stream)
        if ev:
            ev[2].record()
        sharded_phase3(self._dev, xe, B, k, metric, self.n_total, merged.data_ptr(), stream)
        if ev:
            ev[3].record()
            self.phase_events.append(ev)
        return mv

    def search(self, q: torch.Tensor, k: int, metric: str = "cosine", escalate: bool = True):
        """q: (B, d) float64 on this rank's device (identical on every rank).
        Returns (ids (B,k) int64 global rows, dists (B,k) float64, counts (B,) int32, flags (B,)).

        A query whose candidate window could not be proven complete comes back flagged by the device
        path (evdb.h: evdb_store_search_dev).  With `escalate` (the default) no such result leaves
        this call: flagged queries are re-issued on EVERY rank -- the flags are identical everywhere
        (every rank merges the same blobs), so the decision needs no extra collective -- first with a
        256-key window on the scan plan, then through the exhaustive fp64 plan: the ladder
        evdb_store_search_f64 climbs on one GPU (store.cu search_host).  Costs one flag read-back
        per call; escalate=False only enqueues (callers then own the flags)."""
        out = self._search_once(q, k, metric)
        if not escalate:
            return out
        ids, dists, counts, flags = out
        if int(flags.max().item()) == 0:
            return out
        idx = torch.nonzero(flags).flatten()
        self.n_escalations += int(idx.numel())
        for plan, kp_min in (("scan", 256), ("exact", 0)):
            r = self._search_once(q[idx].contiguous(), k, metric, plan=plan, kp_min=kp_min)
            ids[idx] = r[0]; dists[idx] = r[1]; counts[idx] = r[2]; flags[idx] = r[3]
            idx = idx[r[3].to(torch.bool)]
            if idx.numel() == 0:
                break
        return out

    def _search_once(self, q: torch.Tensor, k: int, metric: str, plan: str = "auto", kp_min: int = 0):
        B = q.shape[0]
        forced = plan != "auto" or kp_min != 0
        if not forced and self._two_phase_ok(B, k, metric):
            r = self._search_two_phase(q, k, metric)
            if r is not None:
                return r
        local, gathered, merged, lv, mv, lptrs = self._buffers(B, k, q.device)
        ev = None
        if self.phase_events is not None and q.is_cuda:   # measurement aid (bench.py)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        if self._local_search is not None:      # injected (CPU tests): tensors in, packed here
            if forced:
                ids, dists, counts, flags = self._local_search(q, k, metric, plan=plan, kp_min=kp_min)
            else:
                ids, dists, counts, flags = self._local_search(q, k, metric)
            lv[0].copy_(ids); lv[1].copy_(dists); lv[2].copy_(counts); lv[3].copy_(flags)
        else:
            self._cuda_local_search(q, k, metric, lptrs, plan, kp_min)
        if self.world == 1:
            return lv
        if ev:
            ev[1].record()
        if self._merge is None and self.exchange == "p2p":
            x = self._p2p(B, k, q.device)
            if x is not None:
                stream = _stream_handle(q.device)
                x.push(local.data_ptr(), B, k, stream)      # my blob -> every rank's mailbox, then the epoch flag
                x.merge(B, k, merged.data_ptr(), stream)    # waits on the device for all ranks' flags
                if ev:
                    ev[2].record()
                    self.phase_events.append(ev)
                return mv
        gather_blobs(local, self.world, self.group, out=gathered)
        if self._merge is not None:             # injected (CPU tests)
            g = [blob_views(gathered[r], B, k) for r in range(self.world)]
            out_ids, out_d, out_c = self._merge(torch.stack([x[0] for x in g]), torch.stack([x[1] for x in g]),
                                                torch.stack([x[2] for x in g]), k)
            return out_ids, out_d, out_c, torch.stack([x[3] for x in g]).amax(dim=0)
        self._cuda_merge(gathered, B, k, merged)
        return mv

    def close(self):
        for x in self._xchg.values():
            if x is not None:
                x.close()
        self._xchg = {}
        if self._dev is not None:
            self._dev.close()


class ReplicaGroup:
    """Whole-store replicas, one per GPU, answering DISJOINT query blocks (the reference's own
    scale-out model: cluster_manager replicates whole stores, reference src/cluster_manager.erl:148-171;
    SURVEY 8f-4).  Every per-query cost divides by the number of ranks and no merge is needed -- the
    blocks' results are simply all-gathered -- but every replica reads the WHOLE operand column for its
    query block.  Measured in the same sustained state on B200 (bench.py `layouts`, 1 M x 768, batch 1024)
    it is level with or behind row sharding: N = 2: 1.07 M against 1.22 M QPS, N = 4: 2.03 M against 2.09 M,
    N = 8: 2.7 M against 3.2 M.  Kept as the reference's own model; row sharding is the default."""

    def __init__(self, dtype="f32", device=0, rank=None, world=None, group=None, local_search=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self._local_search = local_search   # injected on CPU (tests); None -> the CUDA library
        self._dev = None if local_search else DeviceStore(dtype=dtype, device=device)
        self._bufs = {}
        self.n_total = 0
        self.n_escalations = 0
        self.lo = self.hi = 0

    def fill_synthetic(self, seed: int, n_total: int, d: int):
        self.n_total = n_total
        self.lo, self.hi = 0, n_total
        if self._dev is not None:
            self._dev.fill_synthetic(seed, n_total, d)

    def bulk_load(self, rows):
        self.n_total = rows.shape[0]
        if self._dev is not None:
            self._dev.bulk_load(rows)

    def search(self, q: torch.Tensor, k: int, metric: str = "cosine", escalate: bool = True):
        """q: (B, d) float64, identical on every rank; rank r answers queries [r*per, (r+1)*per).
        Returns the same tuple as ShardedStore.search, for all B queries, on every rank; flagged
        queries are escalated by the replica that answers them (see ShardedStore.search)."""
        B, d = q.shape
        per = (B + self.world - 1) // self.world
        key = (B, k, q.device)
        if key not in self._bufs:
            w = blob_words(per, k)
            local = torch.zeros((w,), dtype=torch.int64, device=q.device)
            gathered = torch.zeros((self.world, w), dtype=torch.int64, device=q.device)
            self._bufs[key] = (local, gathered, blob_views(local, per, k))
        local, gathered, lv = self._bufs[key]
        b0 = min(B, self.rank * per)
        nb = min(B, b0 + per) - b0
        if nb > 0:
            qb = q[b0:b0 + nb]
            if self._local_search is not None:
                ids, dists, counts, flags = self._local_search(qb, k, metric)
                lv[0][:nb].copy_(ids); lv[1][:nb].copy_(dists); lv[2][:nb].copy_(counts); lv[3][:nb].copy_(flags)
            else:
                self._dev.search_dev(qb.data_ptr(), nb, d, k, metric, 0, lv[0].data_ptr(), lv[1].data_ptr(),
                                     lv[2].data_ptr(), lv[3].data_ptr(), _stream_handle(q.device))
            if escalate and int(lv[3][:nb].max().item()) != 0:
                # a replica owns the whole store: its flagged queries climb the ladder locally
                idx = torch.nonzero(lv[3][:nb]).flatten()
                self.n_escalations += int(idx.numel())
                for plan, kp_min in (("scan", 256), ("exact", 0)):
                    qs = qb[idx].contiguous()
                    if self._local_search is not None:
                        r = self._local_search(qs, k, metric, plan=plan, kp_min=kp_min)
                    else:
                        r = (torch.empty((idx.numel(), k), dtype=torch.int64, device=q.device),
                             torch.empty((idx.numel(), k), dtype=torch.float64, device=q.device),
                             torch.zeros((idx.numel(),), dtype=torch.int32, device=q.device),
                             torch.zeros((idx.numel(),), dtype=torch.int32, device=q.device))
                        self._dev.search_dev_ex(qs.data_ptr(), idx.numel(), d, k, metric, r[0].data_ptr(), r[1].data_ptr(),
                                                r[2].data_ptr(), r[3].data_ptr(), _stream_handle(q.device),
                                                plan=plan, kp_min=kp_min)
                    lv[0][idx] = r[0]; lv[1][idx] = r[1]; lv[2][idx] = r[2]; lv[3][idx] = r[3]
                    idx = idx[r[3].to(torch.bool)]
                    if idx.numel() == 0:
                        break
        if self.world == 1:
            return tuple(t[:B] for t in lv)
        gather_blobs(local, self.world, self.group, out=gathered)
        parts = [blob_views(gathered[r], per, k) for r in range(self.world)]
        return tuple(torch.cat([p_[i] for p_ in parts])[:B] for i in range(4))

    def close(self):
        if self._dev is not None:
            self._dev.close()

