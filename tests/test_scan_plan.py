"""Host logic of the TMA-staged quantized scan (csrc/scan.cu: choose_tma_tile_plan), checked on the
CPU through evdb_debug_scan_tile_plan: the invariants the kernel's mbarrier protocol relies on, the
shared-memory budget, and the plans measured on the B200 (profiles/README.md)."""
import ctypes as C

import pytest

SMS = 148
SMEM_MAX = 227 * 1024


@pytest.fixture(scope="module")
def native():
    """The library loaded WITHOUT a device (no evdb_init): only the host-arithmetic entry is called."""
    from erlvectordb_b200 import _native as N
    N.lib()
    return N


def _plan(native, dtype, dim, window=32, count=12_500_000, sms=SMS):
    out = (C.c_int32 * 8)()
    rc = native.lib().evdb_debug_scan_tile_plan(dtype, dim, window, count, sms, out)
    assert rc in (0, 1), rc
    if rc == 0:
        return None
    keys = ("tpr", "wt", "stages", "stage_bytes", "tile_rows", "rotation", "smem", "groups")
    return dict(zip(keys, list(out)))


def _row_bytes(native, dtype, dim):
    return (dim + 15) // 16 * 16 if dtype == native.U8 else (dim + 31) // 32 * 32 // 2


@pytest.mark.parametrize("window", [32, 64, 128, 512, 1024])
def test_tile_plan_invariants_over_all_shapes(native, window):
    for dtype in (native.U8, native.U4):
        for dim in list(range(1, 300)) + list(range(300, 9000, 37)):
            p = _plan(native, dtype, dim, window)
            rb = _row_bytes(native, dtype, dim)
            if rb < 64:
                assert p is None, (dtype, dim)
                continue
            if p is None:
                continue
            g = p["groups"]
            assert p["wt"] in (1, 2, 4, 8) and g == 8 // p["wt"]
            assert p["tpr"] in (1, 2, 4, 8, 16, 32)
            # a stage always belongs to one consumer group, and every group has a second stage in flight
            assert p["stages"] % g == 0
            assert p["stages"] >= (3 if g == 1 else 2 * g) and p["stages"] <= 8
            # whole tiles: (rows + their 8-byte coefficients) fit the stage; bulk copies are 16-byte multiples
            assert p["tile_rows"] == p["wt"] * (32 // p["tpr"]) * 2 and p["tile_rows"] % 2 == 0
            assert p["stage_bytes"] % 128 == 0
            assert p["tile_rows"] * (rb + 8) <= p["stage_bytes"] < p["tile_rows"] * (rb + 8) + 128
            assert (p["tile_rows"] * rb) % 16 == 0 and (p["tile_rows"] * 8) % 16 == 0
            assert p["tile_rows"] * rb + p["tile_rows"] * 8 < (1 << 20)      # mbarrier tx-count range
            assert p["smem"] + 1024 <= SMEM_MAX
            assert 0 <= p["rotation"] < 8
            assert 12_500_000 // p["tile_rows"] >= 4 * SMS


def test_small_stores_and_tiny_rows_keep_the_register_fed_scan(native):
    assert _plan(native, native.U8, 96, count=100_000) is None          # < 4 tiles per SM
    assert _plan(native, native.U8, 96, count=256 * 4 * SMS - 1) is None
    assert _plan(native, native.U8, 96, count=256 * 4 * SMS) is not None
    assert _plan(native, native.U8, 48) is None                         # 48-byte rows
    assert _plan(native, native.U4, 64) is None                         # 32-byte rows
    assert _plan(native, native.U4, 128) is not None                    # 64-byte rows
    out = (C.c_int32 * 8)()
    assert native.lib().evdb_debug_scan_tile_plan(0, 96, 32, 1000, SMS, out) < 0     # f32 has no such scan
    assert native.lib().evdb_debug_scan_tile_plan(native.U8, 0, 32, 1000, SMS, out) < 0


def test_plans_measured_on_b200(native):
    """The plans behind the numbers in DESIGN.md 5.2 (printed on the GPU box with EVDB_SCAN_DEBUG=1)."""
    p = _plan(native, native.U8, 96)            # BASELINE configs[3] row shape
    assert (p["tpr"], p["wt"], p["stages"], p["stage_bytes"], p["rotation"]) == (2, 8, 4, 26624, 0)
    p = _plan(native, native.U8, 128)           # 128-byte pitch: needs the chunk rotation
    assert (p["tpr"], p["wt"], p["stages"], p["stage_bytes"], p["rotation"]) == (2, 8, 3, 34816, 2)
    p = _plan(native, native.U8, 256, count=4_000_000)
    assert (p["tpr"], p["wt"], p["stages"], p["stage_bytes"], p["rotation"]) == (4, 8, 3, 33792, 4)
    p = _plan(native, native.U4, 1536, count=1_000_000)   # BASELINE configs[4] row shape
    assert (p["tpr"], p["wt"], p["stages"], p["stage_bytes"]) == (8, 4, 4, 24832)
    p = _plan(native, native.U8, 6144, count=300_000)     # 6 KB rows: four consumer groups, one CTA per SM
    assert (p["tpr"], p["wt"], p["stages"], p["stage_bytes"], p["groups"]) == (32, 2, 8, 24704, 4)
    p = _plan(native, native.U8, 96, window=128)          # k = 100: wider candidate lists, shallower ring
    assert (p["wt"], p["stages"]) == (8, 3)
