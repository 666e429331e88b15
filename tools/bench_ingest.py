"""Insert / delete rates through the C ABI on one device (f-3): python tools/bench_ingest.py [dim]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from erlvectordb_b200 import _native as N
from erlvectordb_b200 import synth
from erlvectordb_b200.device_store import DeviceStore

d = int(sys.argv[1]) if len(sys.argv) > 1 else 768
for dtype in ("f32", "u8"):
    st = DeviceStore(dtype=dtype, device=0, capacity_hint=80_000)
    base = np.array(synth.synth(synth.SEED_CORPUS, 0, 4096, d))
    dp = C.POINTER(C.c_double)
    ptrs = [base[i].ctypes.data_as(dp) for i in range(4096)]
    up, dele = N.lib().evdb_store_upsert_f64, N.lib().evdb_store_delete
    for i in range(64):
        up(st.handle, i, ptrs[i], d)
    st.append(base); st.flush()
    c0 = st.stats()["count"]
    n = 20_000
    t0 = time.perf_counter()
    for i in range(n):
        up(st.handle, c0 + i, ptrs[i & 4095], d)
    st.flush()
    up_rate = n / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for i in range(10):
        st.append(base)
    st.flush()
    app = 40960 / (time.perf_counter() - t0)
    cnt = st.stats()["count"]
    moved = C.c_int64()
    t0 = time.perf_counter()
    for i in range(n):
        dele(st.handle, (i * 7919) % (cnt - i), C.byref(moved))
    st.flush()
    de = n / (time.perf_counter() - t0)
    print(f"{dtype} d={d}: {up_rate:,.0f} upserts/s (one row per call), {app:,.0f} rows/s appended 4096 per call, {de:,.0f} deletes/s")
    st.close()
