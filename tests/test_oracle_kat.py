"""Pin the CPU oracle (C and pure-Python twins) against the known-answer vectors derived
from the reference's own fixtures and the committed golden files."""
import json
import os

import numpy as np
import pytest

from conftest import fromhex

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_search_fixture_kat(oracle):
    # test/vector_store_SUITE.erl:66-87: q=[1.0,0.1,0.1] against the 3 basis vectors
    q = [1.0, 0.1, 0.1]
    d1 = oracle.distance(q, [1.0, 0.0, 0.0])
    d2 = oracle.distance(q, [0.0, 1.0, 0.0])
    d3 = oracle.distance(q, [0.0, 0.0, 1.0])
    assert d1 == 0.009852457023325711 and d1.hex() == "0x1.42d86659a7ac0p-7"
    assert d2 == d3 == 0.9009852457023326  # an exact tie, broken by Id
    ids = [b"v1", b"v2", b"v3"]
    rows, dist = oracle.search(np.eye(3), q, 2, ranks=oracle.id_ranks(ids))
    assert [ids[i] for i in rows] == [b"v1", b"v2"]
    # tie-break really follows the Id, not the row
    rows, _ = oracle.search(np.eye(3), q, 2, ranks=oracle.id_ranks([b"v1", b"z", b"a"]))
    assert rows.tolist() == [0, 2]


def test_self_match_kat(oracle):
    assert oracle.distance([1, 2, 3], [1, 2, 3]) == 0.0
    assert oracle.distance([2, 3, 4], [2, 3, 4]) == -2.220446049250313e-16  # negative, not clamped
    assert oracle.distance([1, 0, 0], [1, 0, 0]) == 0.0
    assert oracle.distance([0, 0, 0], [1, 2, 3]) == 1.0  # zero norm clause
    assert oracle.distance([1, 2, 3], [0, 0, 0]) == 1.0


def test_client_demo_kat(oracle):
    q = [1.1, 2.1, 3.1]
    d = {k: oracle.distance(q, v) for k, v in
         {"doc1": [1.0, 2.0, 3.0], "doc2": [2.0, 3.0, 4.0], "doc3": [1.5, 2.5, 3.5]}.items()}
    assert d["doc1"] == 1.4070964634260719e-4
    assert d["doc3"] == 1.5200344460868376e-3
    assert d["doc2"] == 5.5170642681618975e-3


def test_quantizer_kats(oracle):
    c, mn, mx, sc = oracle.quantize_8bit([1.0, 2.5, 3.7, 4.2, 5.9])
    assert c.tolist() == [0, 78, 141, 167, 255] and sc == 0.019215686274509806
    assert oracle.dequantize_8bit(c, mn, sc).tolist() == [1.0, 2.498823529411765, 3.709411764705883,
                                                         4.209019607843137, 5.9]
    c, mn, mx, sc = oracle.quantize_8bit([1.0, 2.0, 3.0])
    assert c.tolist() == [0, 128, 255] and sc == 0.00784313725490196  # 127.5 rounds away from zero
    assert oracle.dequantize_8bit(c, mn, sc)[1] == 2.003921568627451
    p, mn, mx, sc = oracle.quantize_4bit([1.0, 2.0, 3.0, 4.0])
    assert p.tobytes() == bytes([0x05, 0xAF]) and sc == 0.2
    assert oracle.dequantize_4bit(p, 4, mn, sc).tolist() == [1.0, 2.0, 3.0, 4.0]
    v = [float(x) for x in range(1, 51)]
    c, _, _, sc = oracle.quantize_8bit(v)
    assert sc == 0.19215686274509805
    assert c.tolist()[:11] == [0, 5, 10, 16, 21, 26, 31, 36, 42, 47, 52] and c.tolist()[-3:] == [245, 250, 255]
    p, _, _, sc = oracle.quantize_4bit(v)
    assert sc == 3.2666666666666666
    nib = [x for b in p for x in (b >> 4, b & 15)]
    assert nib[:10] == [0, 0, 1, 1, 1, 2, 2, 2, 2, 3] and nib[-3:] == [14, 15, 15]
    with pytest.raises(ArithmeticError):  # Max == Min: badarith in the reference
        oracle.quantize_8bit([2.0, 2.0, 2.0])


def test_python_twin_agrees(oracle):
    rng = np.random.default_rng(7)
    for d in (1, 3, 17, 128):
        a, b = rng.standard_normal(d), rng.standard_normal(d)
        assert oracle.py_cosine_distance(a.tolist(), b.tolist()) == oracle.distance(a, b, "cosine")
        assert oracle.py_euclidean(a.tolist(), b.tolist()) == oracle.distance(a, b, "euclidean")
        assert oracle.py_manhattan(a.tolist(), b.tolist()) == oracle.distance(a, b, "manhattan")
        codes, mn, mx, sc = oracle.py_quantize(a.tolist(), 255) if d > 1 else ([0], 0, 0, 1)
        if d > 1:
            c2, mn2, mx2, sc2 = oracle.quantize_8bit(a)
            assert codes == c2.tolist() and (mn, mx, sc) == (mn2, mx2, sc2)
            c4, _, _, _ = oracle.py_quantize(a.tolist(), 15)
            p4, _, _, _ = oracle.quantize_4bit(a)
            assert oracle.py_pack_4bit(c4) == p4.tobytes()


def test_py_search_matches_c(oracle):
    rng = np.random.default_rng(3)
    rows = rng.integers(-3, 4, size=(40, 5)).astype(np.float64)  # many exact ties
    ids = [bytes([65 + (i * 7) % 26, 48 + i % 10]) + bytes([i]) for i in range(40)]
    q = rng.integers(-3, 4, size=5).astype(np.float64)
    want = oracle.py_search(list(zip(ids, rows.tolist())), q.tolist(), 12)
    got_rows, got_d = oracle.search(rows, q, 12, ranks=oracle.id_ranks(ids))
    assert [ids[i] for i in got_rows] == [i for i, _ in want]
    assert got_d.tolist() == [d for _, d in want]


def test_k_edge_cases(oracle):
    rows = np.eye(3)
    r, d = oracle.search(rows, [1.0, 0.0, 0.0], 10)  # K > N -> all N
    assert len(r) == 3
    r, d = oracle.search(rows, [1.0, 0.0, 0.0], 0)
    assert len(r) == 0
    with pytest.raises(ValueError):
        oracle.search(rows, [1.0, 0.0, 0.0], -1)


def test_bulk_tier_agrees_with_strict(oracle):
    rows = oracle.synth_f64(oracle.SEED_CORPUS, 0, 2000, 96)
    qs = oracle.synth_f64(oracle.SEED_QUERY, 0, 5, 96)
    for metric in ("cosine", "euclidean", "manhattan"):
        bi, bd = oracle.bulk_search(rows, qs, 10, metric)
        for b in range(5):
            si, sd = oracle.search(rows, qs[b], 10, metric)
            assert bi[b].tolist() == si.tolist()
            np.testing.assert_allclose(bd[b], sd, rtol=1e-12, atol=1e-13)
    bi2, bd2 = oracle.bulk_search_synth(oracle.SEED_CORPUS, 2000, 96, qs, 10, "cosine", chunk=512)
    assert bi2.tolist() == oracle.bulk_search(rows, qs, 10, "cosine")[0].tolist()


def test_synth_generator_is_on_the_24bit_grid(oracle):
    v = oracle.synth_f64(oracle.SEED_CORPUS, 0, 64, 32)
    assert np.all(v >= -1.0) and np.all(v < 1.0)
    assert np.all(v * 2 ** 23 == np.round(v * 2 ** 23))
    assert np.array_equal(v.astype(np.float32).astype(np.float64), v)  # exact in fp32
    assert np.array_equal(oracle.synth_f32(oracle.SEED_CORPUS, 0, 64, 32).astype(np.float64), v)
    assert np.array_equal(oracle.synth_f64(oracle.SEED_CORPUS, 10, 5, 32), v[10:15])  # counter based


def test_golden_files_match_oracle(oracle):
    kat = json.load(open(os.path.join(GOLD, "kat.json")))
    sf = kat["search_fixture"]
    for k, v in sf["vectors"].items():
        assert oracle.distance(sf["query"], v) == fromhex(sf["dist"][k])
    cases = json.load(open(os.path.join(GOLD, "synth_small.json")))
    for case in cases:
        rows = oracle.synth_f64(case["seed_corpus"], 0, case["n"], case["d"])
        qs = oracle.synth_f64(case["seed_query"], 0, case["nq"], case["d"])
        assert [x.hex() for x in rows[0, :4]] == case["first_values"]
        for metric in ("cosine", "euclidean", "manhattan"):
            for b, want in enumerate(case["results"][metric]):
                idx, dist = oracle.search(rows, qs[b], case["k"], metric)
                assert idx.tolist() == want["rows"]
                assert [x.hex() for x in dist] == want["dist"]
