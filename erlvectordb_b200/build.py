"""Build libevdb_b200.so (sm_100a only) with nvcc, in-tree.

    python -m erlvectordb_b200.build [--force]

The shared library sits next to this file so that it travels with the repo
snapshot to the GPU box; objects go to build/ (git-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "evdb")
LIB = os.path.join(PKG, "libevdb_b200.so")

SOURCES = ["store.cu", "mstore.cu", "scan.cu", "scan_mq_f32.cu", "scan_mq_bf16.cu", "select.cu", "ingest.cu",
           "gemm_tcgen05.cu", "gemm_i8.cu", "exchange.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libevdb_b200 cannot be built (there is no CPU fallback)")


def _deps() -> list[str]:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(ROOT, "include", "evdb.h"))
    return hdrs


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(nvcc: str, src: str, obj: str) -> None:
    extra = os.environ.get("EVDB_NVCC_EXTRA", "").split()   # e.g. -DEVDB_MW_DEBUG for an instrumented A/B build
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _deps()
    jobs = []
    objs = []
    for name in SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, name.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src, *hdrs]):
            jobs.append((src, obj))
    if jobs:
        if verbose:
            print(f"[evdb build] compiling {len(jobs)} file(s) for sm_100a", file=sys.stderr)
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            for f in [ex.submit(_compile, nvcc, s, o) for s, o in jobs]:
                f.result()
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
