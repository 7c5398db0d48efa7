"""CPU: the rows next to the search path against golden vectors generated
from the UNMODIFIED reference (tests/golden/make_golden.py:golden_scores, golden_kv_store):
the Scorer arithmetic (bigsi/scoring/score.py), the per-window presence strings of score=True
(graph/bigsi.py:232-239) through the oracle, and the v0.3 key/value schema (appendix C)."""
import base64
import re

import numpy as np
import pytest

from bigsi_b200.scoring import Scorer, remove_short_ones, tabulate_score
from oracle import oracle as O
from tests.golden_util import bloom_from_b64, load


def test_scorer_reference_kat():
    # /root/reference/bigsi/tests/scoring.py:10-31: the reference's own known-answer test
    kat = load("scores.json")["reference_kat"]
    got = Scorer(kat["db_size"]).score(kat["s"])
    assert got == kat["expected_by_reference_test"]
    assert got["score"] == 1064.89 and got["mismatches"] == 33 and got["length"] == 1174


def test_scorer_matches_reference_golden():
    cases = load("scores.json")["scorer"]
    assert len(cases) >= 300
    for c in cases:
        got = Scorer(c["db_size"]).score(c["s"])
        want = c["result"]
        assert list(got.keys()) == list(want.keys())
        for key in want:
            a, b = float(got[key]), float(want[key])
            assert a == b or (np.isnan(a) and np.isnan(b)), (key, c["db_size"], c["s"], got[key], want[key])


def test_tabulate_score_counts_like_the_reference():
    # every run but the last is reported one longer than it is (scoring/score.py:19-32)
    assert tabulate_score("0011") == {"0": [3], "1": [2]}
    assert tabulate_score("1") == {"0": [], "1": [1]}
    assert tabulate_score("10") == {"0": [1], "1": [2]}
    assert tabulate_score("110100") == {"0": [2, 2], "1": [3, 2]}
    assert remove_short_ones("11") == "11" and remove_short_ones("1101") == "0001" and remove_short_ones("0111") == "0111"


def _oracle_search_with_score(ix, seq, threshold):
    res = ix.search(seq, threshold)
    if not res:
        return res
    if len(seq) - ix.k + 1 == 1:
        raise IndexError("single window")
    cols = [ix.samples.index(r["sample_name"]) for r in res]
    scorer = Scorer(ix.num_cols)
    for r, s in zip(res, ix.presence_strings(seq, cols)):
        r.update(scorer.score(s))
        r["kmer-presence"] = s
    return res


def test_search_with_score_golden_through_the_oracle():
    for case in load("scores.json")["searches"]:
        blooms = [bloom_from_b64(b) for b in case["blooms_b64"]]
        ix = O.OracleIndex.build(case["k"], case["m"], case["h"], blooms, case["samples"])
        for q in case["queries"]:
            if "raises" in q:
                with pytest.raises(BaseException) as ei:
                    _oracle_search_with_score(ix, q["seq"], q["threshold"])
                assert type(ei.value).__name__ == q["raises"]
            else:
                assert _oracle_search_with_score(ix, q["seq"], q["threshold"]) == q["result"], (q["seq"][:20], q["threshold"])


def kv_of(case):
    return {base64.b64decode(k): base64.b64decode(v) for k, v in case["kv_b64"]}


def metadata_of(kv):
    meta = {"samples": {}, "colours": {}, "colour_count": 0}
    for key, val in kv.items():
        mm = re.fullmatch(rb"metadata:(.*):(int|string)", key)
        if not mm:
            continue
        name = mm.group(1).decode()
        if mm.group(2) == b"string":
            meta["colours"][int(name)] = val.decode()
        elif name == "colour_count":
            meta["colour_count"] = int(val)
        else:
            meta["samples"][name] = int(val)
    return meta


def test_kv_schema_round_trip_through_the_oracle():
    """Rows taken from the reference's store answer the golden queries, and the oracle writes the
    very same key/value pairs back (every key, every byte)."""
    for case in load("kv_store.json"):
        kv = kv_of(case)
        m, n = int(kv[b"number_of_rows:int"]), int(kv[b"number_of_cols:int"])
        assert m == case["m"] and int(kv[b"ksi:num_hashes:int"]) == case["h"] and n == case["num_samples"]
        rb = (n + 7) // 8
        rows = np.stack([np.frombuffer(kv[b"%d:bitarray" % r], dtype=np.uint8)[:rb] for r in range(m)])
        meta = metadata_of(kv)
        samples = [meta["colours"][c] for c in range(n)]
        ix = O.OracleIndex(case["k"], m, case["h"], n, rows=np.ascontiguousarray(rows), samples=samples)
        for q in case["queries"]:
            assert ix.search(q["seq"], q["threshold"]) == q["result"]
        assert ix.to_kv(meta) == kv
