"""Build a second libevdb with extra -D flags for same-box A/B timing (EVDB_LIB_PATH=...).

    python tools/build_variant.py build_ab/libevdb_q3.so -DEVDB_QPLANES=3
    EVDB_LIB_PATH=build_ab/libevdb_q3.so python tools/sweep.py 1000000,1536,u4,cosine,10,1

Objects go next to the output; the in-tree library is not touched.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from erlvectordb_b200 import build as B  # noqa: E402

out = os.path.abspath(sys.argv[1])
defs = sys.argv[2:]
os.makedirs(os.path.dirname(out), exist_ok=True)
nvcc = B._nvcc()
objs = []
for name in B.SOURCES:
    obj = os.path.join(os.path.dirname(out), os.path.basename(out) + "." + name.replace(".cu", ".o"))
    r = subprocess.run([nvcc, *B.NVCC_FLAGS, *defs, "-c", os.path.join(B.CSRC, name), "-o", obj],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(f"nvcc failed for {name}:\n{r.stderr}")
    objs.append(obj)
r = subprocess.run([nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                    "-Xcompiler", "-fPIC", "-Xlinker", "--no-undefined"], capture_output=True, text=True)
if r.returncode != 0:
    sys.exit(f"link failed:\n{r.stderr}")
print(out)
