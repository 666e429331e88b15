"""Row-sharded store: one process per GPU, candidates merged after an NCCL allgather.

The reference has no sharded search (cluster_manager replicates whole stores,
reference src/cluster_manager.erl:148-171); rows are independent and top-k is an
associative merge, so the corpus splits by contiguous row blocks:

    rank r owns global rows [lo_r, hi_r);  every rank sees every query
    local:   evdb_store_search_dev  -> (global id u64, exact fp64 distance) x k
    exchange: torch.distributed all_gather over NCCL/NVLink  (B*k*16 B + B*4 B per rank)
    merge:   evdb_merge_topk_dev on every rank -> identical global top-k everywhere

Because each shard's distances are the exact fp64 values and ids are global rows,
the merged result is bit-identical to the single-GPU result.

torch is plumbing here (device buffers, streams, the process group); the scan,
re-rank and merge are the library's CUDA kernels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _native as N
from .device_store import DeviceStore, merge_topk_dev


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row blocks of ceil(n/world) rows; the last shards may be short or empty."""
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    hi = min(n_total, lo + per)
    return lo, hi


def gather_layout(t: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """all_gather `t` from every rank into a new leading axis: out[g] = rank g's tensor."""
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    if world == 1:
        out[0].copy_(t)
        return out
    if t.is_cuda:  # NCCL: one fused allgather into the [world, ...] buffer
        dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=group)
    else:          # gloo (CPU tests of the host logic)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        for g, p in enumerate(parts):
            out[g].copy_(p)
    return out


def _stream_handle(dev) -> int:
    """cudaStream_t of torch's current stream for the C ABI.  The ABI reads NULL as "the store's
    own stream", so torch's default stream (handle 0) is passed as cudaStreamLegacy (0x1)."""
    h = torch.cuda.current_stream(dev).cuda_stream
    return h if h else 1


class ShardedStore:
    """One shard of a row-sharded store (call from every rank of the process group)."""

    def __init__(self, dtype="f32", device=0, rank=None, world=None, group=None,
                 local_search=None, merge=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.device = device
        self.dtype = dtype
        self._local_search = local_search or self._cuda_local_search
        self._merge = merge or self._cuda_merge
        self._dev = None if local_search else DeviceStore(dtype=dtype, device=device)
        self._bufs = {}
        self.lo = self.hi = 0
        self.n_total = 0
        self.dim = 0

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
    def _cuda_local_search(self, q: torch.Tensor, k: int, metric: str):
        B, d = q.shape
        dev = q.device
        key = (B, k, dev)
        if key not in self._bufs:  # reuse output buffers: no allocator / fill kernels per search
            self._bufs[key] = (torch.empty((B, k), dtype=torch.int64, device=dev),
                               torch.empty((B, k), dtype=torch.float64, device=dev),
                               torch.zeros((B,), dtype=torch.int32, device=dev),
                               torch.zeros((B,), dtype=torch.int32, device=dev))
        ids, dists, counts, flags = self._bufs[key]
        if self.hi > self.lo:
            stream = _stream_handle(dev)
            self._dev.search_dev(q.data_ptr(), B, d, k, metric, self.lo, ids.data_ptr(), dists.data_ptr(),
                                 counts.data_ptr(), flags.data_ptr(), stream)
        return ids, dists, counts, flags

    def _cuda_merge(self, ids, dists, counts, k):
        G, B = counts.shape
        dev = ids.device
        out_ids = torch.empty((B, k), dtype=torch.int64, device=dev)
        out_d = torch.empty((B, k), dtype=torch.float64, device=dev)
        out_c = torch.empty((B,), dtype=torch.int32, device=dev)
        stream = _stream_handle(dev)
        merge_topk_dev(dev.index or 0, ids.data_ptr(), dists.data_ptr(), counts.data_ptr(), G, B, k,
                       out_ids.data_ptr(), out_d.data_ptr(), out_c.data_ptr(), stream)
        return out_ids, out_d, out_c

    def search(self, q: torch.Tensor, k: int, metric: str = "cosine"):
        """q: (B, d) float64 on this rank's device (identical on every rank).
        Returns (ids (B,k) int64 global rows, dists (B,k) float64, counts (B,) int32, flags (B,))."""
        ids, dists, counts, flags = self._local_search(q, k, metric)
        if self.world == 1:
            return ids, dists, counts, flags
        g_ids = gather_layout(ids, self.world, self.group)
        g_d = gather_layout(dists, self.world, self.group)
        g_c = gather_layout(counts, self.world, self.group)
        g_f = gather_layout(flags, self.world, self.group)
        out_ids, out_d, out_c = self._merge(g_ids, g_d, g_c, k)
        return out_ids, out_d, out_c, g_f.amax(dim=0)

    def close(self):
        if self._dev is not None:
            self._dev.close()
