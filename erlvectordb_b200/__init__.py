"""erlvectordb_b200 -- B200-native brute-force kNN engine behind ErlVectorDB's search path.

Layout:
  csrc/                 hand-written sm_100a CUDA kernels + the C ABI (include/evdb.h)
  _native.py            ctypes binding of libevdb_b200.so (fails loudly when missing)
  device_store.py       one device-resident store handle
  vector_store.py       mirror of the reference's vector_store gen_server API
  erlvectordb.py        mirror of the reference facade (create_store/insert/search/...)
  vector_compression.py 8-bit / 4-bit codecs, computed on the device
  sharded.py            row-sharded store over torch.distributed (one process per GPU)
"""
__version__ = "0.1.0"
