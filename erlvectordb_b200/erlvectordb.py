"""Mirror of the reference facade ``erlvectordb`` (reference src/erlvectordb.erl)
for the search path: create_store / insert / search / delete / get_stats, with the
same arities and replies.  ``search/4`` honours the Options map the reference
accepts and ignores (:91-92): ``metric`` selects vector_utils' euclidean /
manhattan forms (additive; the default stays cosine), and ``search_batch`` is the
batch entry point the reference's roadmap lists.
"""
from __future__ import annotations

from . import vector_compression, vector_store

# application:get_env(erlvectordb, Key, Default) -- only the keys this path reads
env = {
    "persistence_enabled": False,     # durability stays with the Erlang side (out of scope here)
    "compression_algorithm": "quantization_8bit",
    "gpu_device": 0,
    "gpu_dtype": "f32",               # f32 | bf16 | quantization_8bit | quantization_4bit
}


def create_store(store_name, options=None):
    """vector_store_sup:start_store/1 (src/erlvectordb.erl:54-55)."""
    if not isinstance(store_name, str):
        raise TypeError("function_clause: store name must be an atom")
    o = dict(options or {})
    return vector_store.start_link(store_name, dtype=o.get("dtype", env["gpu_dtype"]),
                                   device=o.get("device", env["gpu_device"]),
                                   capacity_hint=o.get("capacity_hint", 0))


def delete_store(store_name):
    return vector_store.stop(store_name)


def list_stores():
    return vector_store.which_stores()


def insert(store_name, vector_id, vector, metadata=None):
    """insert/3,4 (:72-77)."""
    return vector_store.insert(store_name, vector_id,
                               {"vector": vector, "metadata": {} if metadata is None else metadata})


def insert_batch(store_name, items):
    """Additive: ``[(Id, Vector, Metadata)]`` in one device call (new ids are appended together)."""
    return vector_store.insert_batch(store_name, [(i, {"vector": v, "metadata": {} if m is None else m})
                                                  for i, v, m in items])


def search(store_name, query_vector, k, options=None):
    """search/3,4 (:88-92)."""
    metric = (options or {}).get("metric", "cosine")
    return vector_store.search(store_name, query_vector, k, metric)


def search_batch(store_name, queries, k, options=None):
    metric = (options or {}).get("metric", "cosine")
    return vector_store.search_batch(store_name, queries, k, metric)


def delete(store_name, vector_id):
    return vector_store.delete(store_name, vector_id)


def get_stats(store_name):
    return vector_store.get_stats(store_name)


def sync(store_name):
    return vector_store.sync(store_name)


def compress_vector(vector, algorithm):
    return vector_compression.compress_vector(vector, algorithm, env["gpu_device"])
