"""Regenerate tests/golden/*.json from the CPU oracle (oracle/evdb_oracle.c).

The reference is pure Erlang and cannot run in this image (no erl/erlc), so the
fixtures are (a) the known-answer values obtained by executing the reference's
exact IEEE-double operation order on its own test inputs
(test/vector_store_SUITE.erl:66-87, test/persistence_SUITE.erl:88-166,
test/compression_SUITE.erl:43-82, examples/mcp_client.py:300-317) and (b) small
synthetic search cases answered by the strict oracle.  Floats are stored as
C99 hex strings so the comparison is bit-exact.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def hx(x):
    return float(x).hex()


def kat():
    q = [1.0, 0.1, 0.1]
    out = {
        "search_fixture": {  # vector_store_SUITE:70-83
            "vectors": {"v1": [1.0, 0.0, 0.0], "v2": [0.0, 1.0, 0.0], "v3": [0.0, 0.0, 1.0]},
            "query": q, "k": 2,
            "dist": {k: hx(O.distance(q, v)) for k, v in
                     {"v1": [1.0, 0.0, 0.0], "v2": [0.0, 1.0, 0.0], "v3": [0.0, 0.0, 1.0]}.items()},
            "expect_ids": ["v1", "v2"],
        },
        "self_match": {  # persistence_SUITE
            "[1,2,3]": hx(O.distance([1, 2, 3], [1, 2, 3])),
            "[2,3,4]": hx(O.distance([2, 3, 4], [2, 3, 4])),
            "[1,0,0]": hx(O.distance([1, 0, 0], [1, 0, 0])),
        },
        "client_demo": {  # examples/mcp_client.py:300-317
            "query": [1.1, 2.1, 3.1],
            "vectors": {"doc1": [1.0, 2.0, 3.0], "doc2": [2.0, 3.0, 4.0], "doc3": [1.5, 2.5, 3.5]},
            "dist": {k: hx(O.distance([1.1, 2.1, 3.1], v)) for k, v in
                     {"doc1": [1.0, 2.0, 3.0], "doc2": [2.0, 3.0, 4.0], "doc3": [1.5, 2.5, 3.5]}.items()},
            "expect_order": ["doc1", "doc3", "doc2"],
        },
    }
    c, mn, mx, sc = O.quantize_8bit([1.0, 2.5, 3.7, 4.2, 5.9])
    out["q8_compression_suite"] = {"vector": [1.0, 2.5, 3.7, 4.2, 5.9], "codes": c.tolist(),
                                   "min": hx(mn), "max": hx(mx), "scale": hx(sc),
                                   "decoded": [hx(x) for x in O.dequantize_8bit(c, mn, sc)]}
    c, mn, mx, sc = O.quantize_8bit([1.0, 2.0, 3.0])
    out["q8_tie"] = {"vector": [1.0, 2.0, 3.0], "codes": c.tolist(), "scale": hx(sc),
                     "decoded": [hx(x) for x in O.dequantize_8bit(c, mn, sc)]}
    p, mn, mx, sc = O.quantize_4bit([1.0, 2.0, 3.0, 4.0])
    out["q4_compression_suite"] = {"vector": [1.0, 2.0, 3.0, 4.0], "packed": p.tobytes().hex(),
                                   "min": hx(mn), "scale": hx(sc),
                                   "decoded": [hx(x) for x in O.dequantize_4bit(p, 4, mn, sc)]}
    v = [float(x) for x in range(1, 51)]
    c, mn, mx, sc = O.quantize_8bit(v)
    out["q8_1_to_50"] = {"codes": c.tolist(), "scale": hx(sc)}
    p, mn, mx, sc = O.quantize_4bit(v)
    out["q4_1_to_50"] = {"packed": p.tobytes().hex(), "scale": hx(sc)}
    p, mn, mx, sc = O.quantize_4bit([0.5, -1.25, 3.0])  # odd length: zero low nibble tail
    out["q4_odd"] = {"vector": [0.5, -1.25, 3.0], "packed": p.tobytes().hex(), "min": hx(mn), "scale": hx(sc)}
    return out


def synth_cases():
    cases = []
    for (n, d, nq, k) in [(512, 64, 8, 10), (300, 96, 4, 7), (257, 33, 4, 5)]:
        rows = O.synth_f64(O.SEED_CORPUS, 0, n, d)
        qs = O.synth_f64(O.SEED_QUERY, 0, nq, d)
        case = {"n": n, "d": d, "nq": nq, "k": k, "seed_corpus": O.SEED_CORPUS, "seed_query": O.SEED_QUERY,
                "first_values": [hx(x) for x in rows[0, :4]], "results": {}}
        for metric in ("cosine", "euclidean", "manhattan"):
            res = []
            for b in range(nq):
                idx, dist = O.search(rows, qs[b], k, metric)
                res.append({"rows": idx.tolist(), "dist": [hx(x) for x in dist]})
            case["results"][metric] = res
        # quantized reload semantics: cosine against Min + c*Scale (vector_persistence.erl:276-284)
        for bits, qf, dq in ((8, O.quantize_8bit, None), (4, O.quantize_4bit, None)):
            deq = np.empty_like(rows)
            codes_hex = []
            for r in range(n):
                c, mn, mx, sc = qf(rows[r])
                deq[r] = O.dequantize_8bit(c, mn, sc) if bits == 8 else O.dequantize_4bit(c, d, mn, sc)
                if r < 4:
                    codes_hex.append({"codes": c.tobytes().hex(), "min": hx(mn), "scale": hx(sc)})
            res = []
            for b in range(nq):
                idx, dist = O.search(deq, qs[b], k, "cosine")
                res.append({"rows": idx.tolist(), "dist": [hx(x) for x in dist]})
            case["results"][f"cosine_q{bits}"] = res
            case[f"codes_q{bits}_first4"] = codes_hex
        cases.append(case)
    return cases


if __name__ == "__main__":
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat(), f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "synth_small.json"), "w") as f:
        json.dump(synth_cases(), f, indent=0, sort_keys=True)
    print("wrote kat.json, synth_small.json")
