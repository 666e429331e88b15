"""DeviceStore: numpy-facing wrapper of one evdb_store handle (one GPU, one store).

This is the level the Erlang NIF shim (erlang/c_src/evdb_nif.c) exposes to the
vector_store gen_server: slots in, (slot, distance) out.  Ids and metadata live
one level up (vector_store.py), exactly as they stay in Erlang state.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class DeviceStore:
    def __init__(self, dtype="f32", dim=0, device=0, capacity_hint=0, gemm_shadow=True, devices=None):
        """devices: a list of CUDA ordinals = ONE store spread over several GPUs of this process behind
        this one handle (evdb_opts.n_shards; the same ordinal may repeat: several shards on one GPU)."""
        L = N.lib()
        self._h = C.c_void_p()
        opts = N.Opts(device=device, dtype=N.DTYPES.get(dtype, dtype), dim=dim,
                      gemm_shadow=1 if gemm_shadow else 0, capacity_hint=capacity_hint)
        if devices is not None and len(devices) > 1:
            opts.n_shards = len(devices)
            for i, dv in enumerate(devices):
                opts.devices[i] = dv
            device = devices[0]
        N.check(L.evdb_store_create(C.byref(opts), C.byref(self._h)), "evdb_store_create")
        self.device = device
        self.dtype = N.DTYPES.get(dtype, dtype)

    # -- lifecycle ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().evdb_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def stats(self) -> dict:
        st = N.Stats()
        N.check(N.lib().evdb_store_stats(self._h, C.byref(st)), "evdb_store_stats")
        return {f: getattr(st, f) for f, _ in N.Stats._fields_}

    def flush(self):
        """Wait for every enqueued upsert / append / delete (they return before the device has run them)."""
        N.check(N.lib().evdb_store_flush(self._h), "evdb_store_flush")

    def set_plan(self, plan):
        N.check(N.lib().evdb_store_set_plan(self._h, N.PLANS.get(plan, plan)), "evdb_store_set_plan")

    def profile(self, enable=True):
        N.check(N.lib().evdb_store_profile(self._h, 1 if enable else 0), "evdb_store_profile")

    def profile_read(self):
        """(samples, total_ms) of the dominant kernel since the last read."""
        n, ms = C.c_int32(0), C.c_double(0.0)
        N.check(N.lib().evdb_store_profile_read(self._h, C.byref(n), C.byref(ms)), "evdb_store_profile_read")
        return n.value, ms.value

    @property
    def count(self) -> int:
        return self.stats()["count"]

    @property
    def dim(self) -> int:
        return self.stats()["dimension"]

    # -- ingest --------------------------------------------------------------
    def upsert(self, slot: int, vec) -> int:
        """Returns the raw status code for DIM_MISMATCH / BAD_VECTOR (the caller maps them
        to the reference's error atoms); raises on anything else."""
        v = np.ascontiguousarray(vec, dtype=np.float64)
        rc = N.lib().evdb_store_upsert_f64(self._h, slot, _p(v, C.c_double), v.shape[0])
        if rc in (N.OK, N.E_DIM_MISMATCH, N.E_BAD_VECTOR):
            return rc
        raise N.EvdbError(rc, "evdb_store_upsert_f64")

    def append(self, rows) -> int:
        """Append n new rows in one call; returns the first slot (status code for the two reference errors)."""
        rows = np.asarray(rows)
        n, d = rows.shape
        first = C.c_uint64(0)
        if rows.dtype == np.float32:
            r = np.ascontiguousarray(rows)
            rc = N.lib().evdb_store_append_f32(self._h, _p(r, C.c_float), n, d, C.byref(first))
        else:
            r = np.ascontiguousarray(rows, dtype=np.float64)
            rc = N.lib().evdb_store_append_f64(self._h, _p(r, C.c_double), n, d, C.byref(first))
        if rc in (N.E_DIM_MISMATCH, N.E_BAD_VECTOR):
            return rc
        N.check(rc, "evdb_store_append")
        return first.value

    def bulk_load(self, rows):
        rows = np.asarray(rows)
        if rows.ndim != 2:
            raise ValueError("rows must be (n, d)")
        n, d = rows.shape
        if rows.dtype == np.float32:
            r = np.ascontiguousarray(rows)
            N.check(N.lib().evdb_store_bulk_load_f32(self._h, _p(r, C.c_float), n, d), "bulk_load_f32")
        else:
            r = np.ascontiguousarray(rows, dtype=np.float64)
            N.check(N.lib().evdb_store_bulk_load_f64(self._h, _p(r, C.c_double), n, d), "bulk_load_f64")

    def bulk_load_codes(self, codes, mins, scales, d):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        mins = np.ascontiguousarray(mins, dtype=np.float64)
        scales = np.ascontiguousarray(scales, dtype=np.float64)
        n = mins.shape[0]
        N.check(N.lib().evdb_store_bulk_load_codes(self._h, _p(codes, C.c_uint8), _p(mins, C.c_double),
                                                   _p(scales, C.c_double), n, d), "bulk_load_codes")

    def fill_synthetic(self, seed, n, d, row0=0):
        N.check(N.lib().evdb_store_fill_synthetic(self._h, seed, row0, n, d), "fill_synthetic")

    def delete(self, slot: int) -> int:
        moved = C.c_int64(-1)
        N.check(N.lib().evdb_store_delete(self._h, slot, C.byref(moved)), "evdb_store_delete")
        return moved.value

    def get(self, slot: int) -> np.ndarray:
        d = self.dim
        out = np.empty(d, dtype=np.float64)
        N.check(N.lib().evdb_store_get_f64(self._h, slot, _p(out, C.c_double), d), "evdb_store_get_f64")
        return out

    def get_codes(self, slot: int):
        d = self.dim
        nb = d if self.dtype == N.U8 else (d + 1) // 2
        codes = np.empty(nb, dtype=np.uint8)
        mn, sc = C.c_double(), C.c_double()
        N.check(N.lib().evdb_store_get_codes(self._h, slot, _p(codes, C.c_uint8), C.byref(mn), C.byref(sc)),
                "evdb_store_get_codes")
        return codes, mn.value, sc.value

    # -- search --------------------------------------------------------------
    def search(self, queries, k: int, metric="cosine"):
        """queries: (B, d) or (d,).  Returns (slots (B,k) u32, dists (B,k) f64, counts (B,))
        or a negative status code for DIM_MISMATCH / BAD_VECTOR."""
        q = np.asarray(queries)
        single = q.ndim == 1
        if single:
            q = q[None, :]
        B, d = q.shape
        m = N.METRICS.get(metric, metric)
        slots = np.empty((B, max(k, 0)), dtype=np.uint32)
        dists = np.empty((B, max(k, 0)), dtype=np.float64)
        counts = np.zeros(B, dtype=np.int32)
        if q.dtype == np.float32:
            q = np.ascontiguousarray(q)
            rc = N.lib().evdb_store_search_f32(self._h, _p(q, C.c_float), B, d, k, m, _p(slots, C.c_uint32),
                                               _p(dists, C.c_double), _p(counts, C.c_int32))
        else:
            q = np.ascontiguousarray(q, dtype=np.float64)
            rc = N.lib().evdb_store_search_f64(self._h, _p(q, C.c_double), B, d, k, m, _p(slots, C.c_uint32),
                                               _p(dists, C.c_double), _p(counts, C.c_int32))
        if rc in (N.E_DIM_MISMATCH, N.E_BAD_VECTOR):
            return rc
        N.check(rc, "evdb_store_search")
        return slots, dists, counts

    def search_dev(self, d_queries_ptr, B, d, k, metric, slot_base, d_ids, d_dists, d_counts, d_flags,
                   stream=0):
        """Raw device-pointer search (enqueue only, no sync)."""
        N.check(N.lib().evdb_store_search_dev(self._h, d_queries_ptr, B, d, k, N.METRICS.get(metric, metric),
                                              slot_base, d_ids, d_dists, d_counts, d_flags, stream),
                "evdb_store_search_dev")


    def search_dev_ex(self, d_queries_ptr, B, d, k, metric, d_ids, d_dists, d_counts, d_flags, stream=0, *,
                      plan="auto", kp_min=0, slot_base=0, slot_stride=1):
        """search_dev with an explicit plan / minimum window (escalation of flagged queries)."""
        o = N.SearchOpts(plan=N.PLANS.get(plan, plan), kp_min=kp_min, slot_base=slot_base, slot_stride=slot_stride)
        N.check(N.lib().evdb_store_search_dev_ex(self._h, d_queries_ptr, B, d, k, N.METRICS.get(metric, metric),
                                                 C.byref(o), d_ids, d_dists, d_counts, d_flags, stream),
                "evdb_store_search_dev_ex")


def gemm_window(k: int, n_total: int) -> int:
    """Candidate window (KP) the tcgen05 plan uses for k results of an n_total-row store (mirrors store.cu)."""
    kk = min(k, n_total)
    want = max(kk + max(kk // 4, 6), 16)
    kp = 1
    while kp < want:
        kp <<= 1
    return 32 if kp <= 32 else (64 if kp <= 64 else (128 if kp <= 128 else kp))


def sharded_phase1(store: "DeviceStore", xw: "Exchange", d_q, B, d, k, metric, slot_base, n_total, stream=0) -> int:
    return N.lib().evdb_store_search_sharded_phase1(store.handle, xw._h, d_q, B, d, k, N.METRICS.get(metric, metric),
                                                    slot_base, n_total, stream)


def sharded_phase2(store, xw, xe, d_q, B, k, metric, n_total, stream=0):
    N.check(N.lib().evdb_store_search_sharded_phase2(store.handle, xw._h, xe._h, d_q, B, k,
                                                     N.METRICS.get(metric, metric), n_total, stream), "sharded_phase2")


def sharded_phase3(store, xe, B, k, metric, n_total, d_out_blob, stream=0):
    N.check(N.lib().evdb_store_search_sharded_phase3(store.handle, xe._h, B, k, N.METRICS.get(metric, metric),
                                                     n_total, d_out_blob, stream), "sharded_phase3")


def merge_topk_dev(device, d_ids, d_dists, d_counts, G, B, k, d_out_ids, d_out_dists, d_out_counts, stream=0):
    N.check(N.lib().evdb_merge_topk_dev(device, d_ids, d_dists, d_counts, G, B, k, d_out_ids, d_out_dists,
                                        d_out_counts, stream), "evdb_merge_topk_dev")


def merge_topk_packed_dev(device, d_blobs, G, B, k, d_out_blob, stream=0):
    N.check(N.lib().evdb_merge_topk_packed_dev(device, d_blobs, G, B, k, d_out_blob, stream),
            "evdb_merge_topk_packed_dev")


class Exchange:
    """Peer-memory exchange of packed result blobs (evdb_exchange_*): one object per rank."""

    def __init__(self, device: int, rank: int, world: int, max_words: int):
        self._h = C.c_void_p()
        self.handle_bytes = (C.c_uint8 * 64)()
        N.check(N.lib().evdb_exchange_create(device, rank, world, max_words, C.byref(self._h),
                                             C.cast(self.handle_bytes, C.c_void_p)), "evdb_exchange_create")
        self.world, self.max_words = world, max_words

    @property
    def mailbox(self) -> int:
        return N.lib().evdb_exchange_mailbox(self._h)

    def connect(self, all_handles: bytes):
        buf = (C.c_uint8 * len(all_handles)).from_buffer_copy(all_handles)
        N.check(N.lib().evdb_exchange_connect(self._h, C.cast(buf, C.c_void_p)), "evdb_exchange_connect")

    def connect_ptrs(self, mailboxes):
        arr = (C.c_void_p * len(mailboxes))(*mailboxes)
        N.check(N.lib().evdb_exchange_connect_ptrs(self._h, arr), "evdb_exchange_connect_ptrs")

    def push(self, d_blob: int, B: int, k: int, stream=0):
        N.check(N.lib().evdb_exchange_push(self._h, d_blob, B, k, stream), "evdb_exchange_push")

    def merge(self, B: int, k: int, d_out_blob: int, stream=0):
        N.check(N.lib().evdb_exchange_merge(self._h, B, k, d_out_blob, stream), "evdb_exchange_merge")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().evdb_exchange_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
