"""Small end-to-end exercise of every kernel family (run under compute-sanitizer)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from erlvectordb_b200 import synth
from erlvectordb_b200.device_store import DeviceStore

for dtype, n, d, B, k, metrics in [("f32", 20003, 100, 40, 10, ("cosine", "euclidean", "manhattan")),
                                   ("f32", 9000, 128, 136, 100, ("cosine", "euclidean")),
                                   ("bf16", 5000, 72, 3, 5, ("cosine",)),
                                   ("u8", 6000, 96, 2, 10, ("cosine", "euclidean")),
                                   ("u4", 6000, 70, 1, 10, ("cosine",))]:
    st = DeviceStore(dtype=dtype, device=0)
    st.fill_synthetic(synth.SEED_CORPUS, n, d)
    q = synth.synth(synth.SEED_QUERY, 0, B, d)
    for m in metrics:
        r = st.search(q, k, m)
        r1 = st.search(q[:1], k, m)
        assert r[0][0].tolist() == r1[0][0].tolist(), (dtype, m)
        print(dtype, n, d, B, k, m, "plan", st.stats()["last_plan"], "ok", flush=True)
    if dtype == "f32":
        v = np.ones(d)
        assert st.upsert(n, v) == 0 and st.upsert(3, v * 2) == 0
        st.delete(0)
        st.search(q, k, "euclidean")
        st.append(np.random.default_rng(0).standard_normal((50, d)))
        st.search(q, k, "cosine")
    st.close()
print("SANITY_OK")
