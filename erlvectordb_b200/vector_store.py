"""Host-side mirror of the reference's ``vector_store`` gen_server
(reference src/vector_store.erl), backed by a device-resident store.

Same names, argument meaning and replies as the Erlang module so that tests read
like test/vector_store_SUITE.erl: ``{ok, X}`` is ``("ok", X)``, a bare ``ok`` is
``"ok"``, ``{error, Reason}`` is ``("error", "reason")``.  One registered store
per name; a per-store lock plays the gen_server mailbox (operations on one
store are totally ordered, different stores run concurrently).

What lives where (as in the NIF deployment):
  device (libevdb_b200):  packed vectors, distance scan, top-k, exact re-rank
  this process (Erlang state in production): Id <-> slot maps, metadata, the
  final ``{Distance, Id}`` term-order tie-break of lists:sort/1.
"""
from __future__ import annotations

import math
import threading
from numbers import Integral, Real

import numpy as np

from . import _native as N
from .device_store import DeviceStore

_registry: dict[str, "VectorStore"] = {}
_registry_lock = threading.Lock()


class FunctionClause(Exception):
    """The reference's store process crashes here (e.g. lists:sublist/2 with K < 0)."""


def term_key(t):
    """Erlang term order, enough for Ids: number < atom(str) < tuple < list < binary(bytes)."""
    if isinstance(t, bool):
        return (1, str(t).lower())
    if isinstance(t, Real):
        return (0, t)
    if isinstance(t, str):
        return (1, t)
    if isinstance(t, tuple):
        return (2, len(t), tuple(term_key(x) for x in t))
    if isinstance(t, list):
        return (3, tuple(term_key(x) for x in t))
    if isinstance(t, (bytes, bytearray)):
        return (4, bytes(t))
    return (5, repr(t))


def validate_vector(vector, dimension):
    """validate_vector/2 (src/vector_store.erl:213-225)."""
    if isinstance(vector, np.ndarray):
        if vector.ndim != 1 or vector.dtype.kind not in "fiu":
            return ("error", "invalid_vector_format")
        n = vector.shape[0]
    elif isinstance(vector, (list, tuple)):
        if not all(isinstance(x, Real) and not isinstance(x, bool) for x in vector):
            return ("error", "invalid_vector_format")
        n = len(vector)
    else:
        return ("error", "invalid_vector_format")
    if dimension is None or n == dimension:
        return ("ok", n)
    return ("error", "dimension_mismatch")


class VectorStore:
    def __init__(self, name, dtype="f32", device=0, capacity_hint=0, gemm_shadow=True):
        self.name = name
        self.dimension = None
        self._dev = DeviceStore(dtype=dtype, dim=0, device=device, capacity_hint=capacity_hint,
                                gemm_shadow=gemm_shadow)
        self._id2slot: dict = {}
        self._slot2id: list = []
        self._meta: dict = {}
        self._ordered = True       # slot order == Id term order (no tie-break widening needed)
        self._lock = threading.Lock()
        self.persistence_enabled = False

    # -- handle_call({insert, Id, #{vector, metadata}}) :113-141 ----------------
    def insert(self, vector_id, vector_data):
        with self._lock:
            vector, metadata = vector_data["vector"], vector_data["metadata"]
            v = validate_vector(vector, self.dimension)
            if v[0] == "error":
                return v
            try:
                arr = np.asarray(vector, dtype=np.float64)
            except (OverflowError, ValueError):
                return ("error", "invalid_vector_format")
            if arr.size == 0 or not np.all(np.isfinite(arr)):
                return ("error", "invalid_vector_format")
            slot = self._id2slot.get(vector_id)
            new = slot is None
            if new:
                slot = len(self._slot2id)
            rc = self._dev.upsert(slot, arr)
            if rc == N.E_DIM_MISMATCH:
                return ("error", "dimension_mismatch")
            if rc == N.E_BAD_VECTOR:
                return ("error", "invalid_vector_format")
            if new:
                if self._slot2id and self._ordered and not term_key(self._slot2id[-1]) < term_key(vector_id):
                    self._ordered = False
                self._id2slot[vector_id] = slot
                self._slot2id.append(vector_id)
            self._meta[vector_id] = metadata
            self.dimension = v[1]
            return "ok"

    def insert_batch(self, items):
        """Additive API: ``[(Id, #{vector, metadata})]`` inserted as ONE device call when every Id
        is new (evdb_store_append_*); equals calling ``insert`` on each in order.  Stops at the first
        invalid item with the reference's error tuple (items before it are kept, as N calls would)."""
        pending = []          # new ids waiting for one append
        seen = set()

        def flush():
            if not pending:
                return "ok"
            rows = np.stack([p[2] for p in pending])
            with self._lock:
                first = self._dev.append(rows)
                if first == N.E_DIM_MISMATCH:
                    return ("error", "dimension_mismatch")
                if first == N.E_BAD_VECTOR:
                    return ("error", "invalid_vector_format")
                for j, (vid, meta, _) in enumerate(pending):
                    if self._slot2id and self._ordered and not term_key(self._slot2id[-1]) < term_key(vid):
                        self._ordered = False
                    self._id2slot[vid] = first + j
                    self._slot2id.append(vid)
                    self._meta[vid] = meta
                self.dimension = rows.shape[1]
            pending.clear()
            seen.clear()
            return "ok"

        for vid, data in items:
            dim = self.dimension if self.dimension is not None else (pending[0][2].shape[0] if pending else None)
            v = validate_vector(data["vector"], dim)
            arr = None
            if v[0] == "ok":
                try:
                    arr = np.asarray(data["vector"], dtype=np.float64)
                except (OverflowError, ValueError):
                    arr = None
            if arr is None or arr.size == 0 or not np.all(np.isfinite(arr)):
                r = flush()
                return r if r != "ok" else (v if v[0] == "error" else ("error", "invalid_vector_format"))
            if vid in self._id2slot or vid in seen:   # an overwrite: keep call order, one at a time
                r = flush()
                if r != "ok":
                    return r
                r = self.insert(vid, data)
                if r != "ok":
                    return r
                continue
            pending.append((vid, data["metadata"], arr))
            seen.add(vid)
        return flush()

    # -- handle_call({search, Q, K}) :143-150, perform_search/3 :227-236 --------
    def search(self, query_vector, k, metric="cosine"):
        r = self.search_batch([query_vector], k, metric)
        if r[0] == "error":
            return r
        return ("ok", r[1][0])

    def search_batch(self, queries, k, metric="cosine"):
        """B searches in one device call; element b equals ``search(queries[b], k)``."""
        with self._lock:
            for q in queries:
                v = validate_vector(q, self.dimension)
                if v[0] == "error":
                    return v
            if not isinstance(k, Integral) or isinstance(k, bool) or k < 0:
                raise FunctionClause("lists:sublist/2")  # the reference store crashes
            B = len(queries)
            n = len(self._slot2id)
            if n == 0 or k == 0 or B == 0:
                return ("ok", [[] for _ in range(B)])
            try:
                q = np.asarray(queries, dtype=np.float64)
            except (OverflowError, ValueError):
                return ("error", "invalid_vector_format")
            if q.ndim != 2 or not np.all(np.isfinite(q)):
                return ("error", "invalid_vector_format")
            kk = min(k, n)
            k2 = kk if self._ordered else min(n, kk + 16)
            while True:
                res = self._dev.search(q, k2, metric)
                if isinstance(res, int):
                    return ("error", "dimension_mismatch" if res == N.E_DIM_MISMATCH
                            else "invalid_vector_format")
                slots, dists, counts = res
                if self._ordered or k2 >= n:
                    break
                # an exact-distance tie group straddling the k2 window could hide Ids that sort
                # first: widen until the window ends outside the group of the kk-th result
                if np.any(dists[:, kk - 1] == dists[:, k2 - 1]):
                    k2 = min(n, k2 * 2)
                    continue
                break
            out = []
            for b in range(B):
                c = int(counts[b])
                rows = [(float(dists[b, j]), self._slot2id[int(slots[b, j])]) for j in range(c)]
                if not self._ordered:
                    rows.sort(key=lambda t: (t[0], term_key(t[1])))  # lists:sort/1 on {Distance, Id, _}
                out.append([(i, self._meta[i], d) for d, i in rows[:kk]])
            return ("ok", out)

    # -- handle_call({delete, Id}) :152-164 -------------------------------------
    def delete(self, vector_id):
        with self._lock:
            slot = self._id2slot.pop(vector_id, None)
            if slot is None:
                return "ok"  # maps:remove of a missing key is a no-op
            self._meta.pop(vector_id, None)
            moved = self._dev.delete(slot)
            last_id = self._slot2id.pop()
            if moved >= 0:
                self._slot2id[slot] = last_id
                self._id2slot[last_id] = slot
                self._ordered = False
            return "ok"

    # -- handle_call(get_stats) :166-173 ------------------------------------------
    def get_stats(self):
        with self._lock:
            st = {"name": self.name, "count": len(self._slot2id), "dimension": self.dimension,
                  "persistence_enabled": self.persistence_enabled}
            st["gpu"] = self._dev.stats()  # additive keys; the reference's four stay as they are
            return ("ok", st)

    # -- handle_call(sync) :175-182 ---------------------------------------------------
    def sync(self):
        return ("error", "persistence_disabled")

    # -- handle_call(get_all_vectors) :184-190 ------------------------------------------
    def get_all_vectors(self):
        with self._lock:
            return ("ok", {i: {"vector": self._dev.get(s).tolist(), "metadata": self._meta[i]}
                           for i, s in self._id2slot.items()})

    # -- init/1 bulk load :68-96 ------------------------------------------------------------
    def load_vectors(self, loaded: dict):
        """``loaded``: Id -> #{vector, metadata} as vector_persistence:load_vectors/1 returns."""
        with self._lock:
            ids = list(loaded.keys())
            if not ids:
                return "ok"
            rows = np.asarray([loaded[i]["vector"] for i in ids], dtype=np.float64)
            self._dev.bulk_load(rows)
            self._slot2id = ids
            self._id2slot = {i: s for s, i in enumerate(ids)}
            self._meta = {i: loaded[i]["metadata"] for i in ids}
            self.dimension = rows.shape[1]
            keys = [term_key(i) for i in ids]
            self._ordered = all(keys[j] < keys[j + 1] for j in range(len(keys) - 1))
            return "ok"

    def load_compressed(self, records: dict):
        """Compressed records (vector_compression maps) straight to device code columns --
        what init/1 sees after decompress_if_needed (vector_persistence.erl:276-284)."""
        with self._lock:
            ids = list(records.keys())
            if not ids:
                return "ok"
            first = records[ids[0]]["vector"]
            d = first["metadata"].get("length", len(first["data"]))
            codes = np.frombuffer(b"".join(records[i]["vector"]["data"] for i in ids), dtype=np.uint8)
            mins = np.array([records[i]["vector"]["metadata"]["min"] for i in ids], dtype=np.float64)
            scales = np.array([records[i]["vector"]["metadata"]["scale"] for i in ids], dtype=np.float64)
            self._dev.bulk_load_codes(codes, mins, scales, d)
            self._slot2id = ids
            self._id2slot = {i: s for s, i in enumerate(ids)}
            self._meta = {i: records[i]["metadata"] for i in ids}
            self.dimension = d
            keys = [term_key(i) for i in ids]
            self._ordered = all(keys[j] < keys[j + 1] for j in range(len(keys) - 1))
            return "ok"

    # -- terminate/2 :201-207 -----------------------------------------------------------------
    def terminate(self):
        with self._lock:
            self._dev.close()


# ---- module API (same arity and names as the Erlang exports, :16-19,38-57) --------
def start_link(name, **opts):
    with _registry_lock:
        if name in _registry:
            return ("error", ("already_started", _registry[name]))
        s = VectorStore(name, **opts)
        _registry[name] = s
        return ("ok", s)


def stop(name):
    with _registry_lock:
        s = _registry.pop(name, None)
    if s is None:
        return ("error", "not_found")
    s.terminate()
    return "ok"


def _whereis(name) -> VectorStore:
    s = _registry.get(name)
    if s is None:
        raise LookupError(f"noproc: {name}")  # gen_server:call to an unregistered name exits
    return s


def insert(store_name, vector_id, vector_data):
    return _whereis(store_name).insert(vector_id, vector_data)


def search(store_name, query_vector, k, metric="cosine"):
    return _whereis(store_name).search(query_vector, k, metric)


def search_batch(store_name, queries, k, metric="cosine"):
    return _whereis(store_name).search_batch(queries, k, metric)


def insert_batch(store_name, items):
    return _whereis(store_name).insert_batch(items)


def delete(store_name, vector_id):
    return _whereis(store_name).delete(vector_id)


def get_stats(store_name):
    return _whereis(store_name).get_stats()


def sync(store_name):
    return _whereis(store_name).sync()


def get_all_vectors(store_name):
    return _whereis(store_name).get_all_vectors()


def which_stores():
    return list(_registry.keys())
