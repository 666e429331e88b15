"""The built library really contains the Blackwell paths it claims (B200_PROFILING.md: the SASS
mnemonics that prove tcgen05 / TMEM / TMA), checked here on the CPU with cuobjdump -- a build that
silently lost them (a flag change, a refactor that fell back to generic loads) fails the suite
before any GPU time is spent."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "erlvectordb_b200", "libevdb_b200.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    from erlvectordb_b200 import build
    build.build()
    txt = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = {}
    name = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    assert funcs, "no device code in the library"
    assert "sm_100a" in subprocess.run([exe, "-lelf", LIB], capture_output=True, text=True).stdout
    return {k: "\n".join(v) for k, v in funcs.items()}


def _pick(sass, *needles):
    out = {k: v for k, v in sass.items() if all(n in k for n in needles)}
    assert out, f"no kernel matching {needles}"
    return out


def test_gemm_kernel_uses_tcgen05_tmem_and_tma(sass):
    one = _pick(sass, "gemm_topk_kernelILb0E")
    pair = _pick(sass, "gemm_topk_kernelILb1E")
    for body in list(one.values()) + list(pair.values()):
        assert "UTCHMMA" in body          # tcgen05.mma
        assert "LDTM.x32" in body         # tcgen05.ld 32x32b.x32 (TMEM -> registers)
        assert "UTMALDG.2D" in body       # cp.async.bulk.tensor.2d
        assert "UTCBAR" in body           # tcgen05.commit -> mbarrier
        assert "FMNMX3" in body           # the epilogue's 3-input max tree
        assert "HMMA" not in body.replace("UTCHMMA", "")   # no mma.sync anywhere
    for body in pair.values():
        assert "UTCHMMA.2CTA" in body and "UTMALDG.2D.2CTA" in body and "UTCBAR.2CTA.MULTICAST" in body
    for body in one.values():
        assert "UTCHMMA.2CTA" not in body


def test_i8_kernel_uses_integer_tcgen05_and_shared_space_loads(sass):
    ks = _pick(sass, "gemm_i8_topk_kernel")
    assert len(ks) == 8                   # {mantissa-trick, I2F} x {resident, streamed query planes} x {8, 16 epilogue warps}
    for name, body in ks.items():
        assert "UTCIMMA" in body, name    # tcgen05.mma.kind::i8
        assert "LDTM.x32" in body and "UTMALDG.2D" in body and "UTCBAR" in body, name
        assert "LDS.64" in body, name     # row coefficients: shared-space loads, not generic ones
        if "ILb1E" in name:
            assert "LDS.128" in body, name  # the coarse filter's {256 cx, K', cy, cx}
        assert "UTCHMMA" not in body, name


def test_tma_staged_scan_uses_bulk_copies_mbarriers_and_dp4a(sass):
    ks = _pick(sass, "scan_quant_tma_kernel")
    assert len(ks) == 12                  # {u8, u4} x six lanes-per-row variants
    for name, body in ks.items():
        assert body.count("UBLKCP.S.G") == 2, name        # codes + coefficients of a tile
        assert "SYNCS.ARRIVE.TRANS64" in body and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in body
        assert "IDP.4A.S8.U8" in body and "IDP.4A.U8.U8" in body
        assert "LDS.128" in body          # rows and query digits come out of shared memory


def test_scan_kernels_stream_with_128_bit_loads(sass):
    for name, body in _pick(sass, "scan_float_kernel").items():
        assert "LDG.E.NA.128" in body or "LDG.E.128.NA" in body or re.search(r"LDG\.E\.[A-Z.]*128", body), name
    for name, body in _pick(sass, "scan_quant_kernelI").items():
        assert "IDP.4A" in body and re.search(r"LDG\.E\.[A-Z.]*128", body), name


def test_exact_rerank_keeps_products_and_sums_separate(sass):
    """The re-rank follows the reference's operation order: products and sums are rounded
    separately (DMUL then DADD chains; __dmul_rn/__dadd_rn forbid contraction).  DFMA may only
    appear in the correctly rounded division / square-root sequences, i.e. far fewer than DADDs."""
    for needle in ("select_warp_kernel", "select_kernelILi0ELi1024E", "shard_rerank_kernel"):
        for name, body in _pick(sass, needle).items():
            n_add, n_mul, n_fma = body.count(" DADD "), body.count(" DMUL "), body.count(" DFMA ")
            assert n_add >= 16 and n_mul >= 4, name
            assert n_fma < n_add, (name, n_fma, n_add)
