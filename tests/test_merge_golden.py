"""BIGSI.merge (bigsi/graph/bigsi.py:252-260) against stores the unmodified reference wrote after merging
(tests/golden/make_golden.py:golden_merge): bit-concatenated rows, sample renaming, queries."""
import numpy as np
import pytest

from bigsi_b200.bigsi import merge_packed_rows
from bigsi_b200.metadata import SampleMetadata
from oracle import oracle as O
from tests.golden_util import load
from tests.test_scoring_kv_golden import kv_of, metadata_of


def _oracle_index(case, seqs, names):
    k, m, h = case["k"], case["m"], case["h"]
    blooms = [O.OracleIndex.bloom(k, m, h, [O.canonical(x) for x in O.seq_to_kmers(s, k)]) for s in seqs]
    return O.OracleIndex.build(k, m, h, blooms, names)


def test_merge_rows_and_metadata_on_cpu():
    for case in load("merge.json"):
        kv = kv_of(case)
        a = _oracle_index(case, case["seqs1"], case["names1"])
        b = _oracle_index(case, case["seqs2"], case["names2"])
        n1, n2 = len(case["names1"]), len(case["names2"])
        rows = merge_packed_rows(a.rows, n1, b.rows, n2)
        want = np.stack([np.frombuffer(kv[b"%d:bitarray" % r], dtype=np.uint8) for r in range(case["m"])])
        assert np.array_equal(rows, want)
        # sample bookkeeping of graph/metadata.py:74-80 on the host-side SampleMetadata
        store, store2 = {}, {}
        sm, sm2 = SampleMetadata(store), SampleMetadata(store2)
        sm.add_samples(case["names1"])
        sm2.add_samples(case["names2"])
        if case["deleted_in_2"]:
            sm2.delete_sample(case["deleted_in_2"])
        for c in range(sm2.num_samples):
            s = sm2.colour_to_sample(c)
            try:
                sm.add_sample(s)
            except ValueError:
                sm.add_sample(s + "_duplicate_in_merge")
        meta = metadata_of(kv)
        assert sm.num_samples == case["num_samples"] == meta["colour_count"]
        assert {c: sm.colour_to_sample(c) for c in range(sm.num_samples)} == meta["colours"]
        merged = O.OracleIndex(case["k"], case["m"], case["h"], n1 + n2, rows=np.ascontiguousarray(rows),
                               samples=[meta["colours"][c] for c in range(n1 + n2)])
        for q in case["queries"]:
            assert merged.search(q["seq"], q["threshold"]) == q["result"]


def test_merge_packed_rows_edges():
    rng = np.random.default_rng(0)
    for n1, n2 in ((0, 5), (5, 0), (8, 8), (7, 1), (1, 7), (13, 22)):
        A = rng.integers(0, 2, size=(9, n1), dtype=np.uint8)
        Bm = rng.integers(0, 2, size=(9, n2), dtype=np.uint8)
        pa = np.packbits(A, axis=1) if n1 else np.zeros((9, 0), dtype=np.uint8)
        pb = np.packbits(Bm, axis=1) if n2 else np.zeros((9, 0), dtype=np.uint8)
        got = merge_packed_rows(pa, n1, pb, n2)
        assert np.array_equal(got, np.packbits(np.concatenate([A, Bm], axis=1), axis=1))


@pytest.mark.gpu
def test_merge_matches_reference_store():
    import bigsi_b200 as B

    for ci, case in enumerate(load("merge.json")):
        k, m, h = case["k"], case["m"], case["h"]
        cfg1 = {"k": k, "m": m, "h": h, "storage-config": {"filename": "merge-a-%d" % ci}}
        cfg2 = {"k": k, "m": m, "h": h, "storage-config": {"filename": "merge-b-%d" % ci}}
        b1 = B.BIGSI.build(cfg1, [B.BIGSI.bloom(cfg1, B.seq_to_kmers(s, k)) for s in case["seqs1"]], case["names1"])
        b2 = B.BIGSI.build(cfg2, [B.BIGSI.bloom(cfg2, B.seq_to_kmers(s, k)) for s in case["seqs2"]], case["names2"])
        try:
            if case["deleted_in_2"]:
                b2.delete_sample(case["deleted_in_2"])
            b1.merge(b2)
            assert b1.num_samples == case["num_samples"] and b1.index.num_cols == case["num_samples"]
            merged = B.BIGSI(cfg1)  # a fresh handle on the resident store sees the merged index
            for q in case["queries"]:
                assert merged.search(q["seq"], q["threshold"]) == q["result"]
            assert merged.to_kv() == kv_of(case)
        finally:
            b1.delete()
            b2.delete()
