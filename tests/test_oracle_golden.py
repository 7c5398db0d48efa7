"""CPU: pins the oracle (oracle/bigsi_oracle.c + oracle/oracle.py) against golden vectors
generated from the UNMODIFIED reference (tests/golden/make_golden.py) and the reference's
own known-answer tests."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.golden_util import bloom_from_b64, load, rows_from_b64


def test_reference_kat_generate_hashes():
    # /root/reference/bigsi/tests/bloom/test_create_bloomfilter.py:5-8
    assert O.generate_hashes("ATT", 3, 25) == {2, 15, 17}
    assert O.generate_hashes("ATT", 1, 25) == {15}
    assert O.generate_hashes("ATT", 2, 50) == {15, 27}
    for kat in load("hashes.json")["kat"]:
        assert sorted(O.generate_hashes(kat["element"], kat["h"], kat["m"])) == kat["set"]


def test_canonical_golden():
    for kmer, can in load("hashes.json")["canonical"]:
        assert O.canonical(kmer) == can, kmer


def test_hash_rows_golden():
    cases = load("hashes.json")["cases"]
    for c in cases:
        got = O.hash_kmers([c["kmer"]], len(c["kmer"]), c["h"], c["m"])[0].tolist()
        assert got == c["rows"], c


def _index_from_case(case, rows_b64=None, row_bytes=None, samples=None):
    samples = samples or case["samples"]
    rows = rows_from_b64(rows_b64 or case["rows_b64"], case["m"], row_bytes or case["row_bytes"])
    return O.OracleIndex(case["k"], case["m"], case["h"], len(samples), rows=rows, samples=samples)


def _check_queries(ix, queries):
    for q in queries:
        if "raises" in q:
            with pytest.raises(BaseException) as ei:
                ix.search(q["seq"], q["threshold"])
            assert type(ei.value).__name__ == q["raises"]
        else:
            assert ix.search(q["seq"], q["threshold"]) == q["result"], (q["seq"], q["threshold"])


def test_search_golden_on_reference_rows():
    g = load("search_cases.json")
    for case in g["cases"]:
        ix = _index_from_case(case)
        _check_queries(ix, case["queries"])
        assert ix.lookup(case["lookup_in"]) == case["lookup"]


def test_build_matches_reference_rows():
    """oracle build (bloom + transpose) reproduces the reference's stored row bytes."""
    g = load("search_cases.json")
    for case in g["cases"]:
        k, m, h = case["k"], case["m"], case["h"]
        blooms = []
        for seq, ref_b64 in zip(case["sample_seqs"], case["blooms_b64"]):
            b = O.OracleIndex.bloom(k, m, h, [O.canonical(x) for x in O.seq_to_kmers(seq, k)])
            assert np.array_equal(b, bloom_from_b64(ref_b64))
            blooms.append(b)
        ix = O.OracleIndex.build(k, m, h, blooms, case["samples"])
        assert np.array_equal(ix.rows, rows_from_b64(case["rows_b64"], m, case["row_bytes"]))


def test_insert_and_delete_golden():
    g = load("search_cases.json")
    n_ins = n_del = 0
    for case in g["cases"]:
        samples = list(case["samples"])
        if "insert" in case:
            ins = case["insert"]
            samples = samples + [ins["sample"]]
            ix = _index_from_case(case, ins["rows_b64"], ins["row_bytes"], samples)
            assert ix.num_cols == ins["num_samples"]
            _check_queries(ix, ins["queries"])
            n_ins += 1
        if "delete" in case:
            d = case["delete"]
            samples = [O.DELETION_SPECIAL_SAMPLE_NAME if s == d["sample"] else s for s in samples]
            rows_b64 = case["insert"]["rows_b64"] if "insert" in case else case["rows_b64"]
            row_bytes = case["insert"]["row_bytes"] if "insert" in case else case["row_bytes"]
            ix = _index_from_case(case, rows_b64, row_bytes, samples)
            _check_queries(ix, d["queries"])
            n_del += 1
    assert n_ins >= 3 and n_del >= 2


def test_reference_end_to_end_kat():
    # /root/reference/bigsi/tests/graph/test_end_to_end.py:69-131
    kat = load("search_cases.json")["reference_kat"]
    k, m, h = 3, 1000, 3

    def mk(seqs, names):
        blooms = [O.OracleIndex.bloom(k, m, h, [O.canonical(x) for x in O.seq_to_kmers(s, k)]) for s in seqs]
        return O.OracleIndex.build(k, m, h, blooms, names)

    ix = mk(["ATACACAAT", "ACAGAGAAC"], ["a", "b"])
    assert ix.search("ATACACAAT")[0] == {"percent_kmers_found": 100, "num_kmers": 6, "num_kmers_found": 6, "sample_name": "a"}
    assert ix.search("ACAGAGAAC")[0] == {"percent_kmers_found": 100, "num_kmers": 6, "num_kmers_found": 6, "sample_name": "b"}
    assert ix.search("ACAGTTAAC") == []
    _check_queries(ix, kat["exact"])
    ix = mk(["ATACACAAT", "ATACACAAC"], ["a", "b"])
    _check_queries(ix, kat["inexact"])
    assert ix.lookup("AAT") == kat["inexact_lookup"] == {"AAT": "10"}
    res = ix.search("ATACACAAT", 0.5)
    assert res[1] == {"percent_kmers_found": 83.33, "num_kmers": 6, "num_kmers_found": 5, "sample_name": "b"}


def test_config1_golden():
    c = load("config1.json")
    ix = _index_from_case(c)
    _check_queries(ix, c["queries"])
    blooms = [O.OracleIndex.bloom(c["k"], c["m"], c["h"], [O.canonical(x) for x in km]) for km in c["sample_kmers"]]
    assert [int(np.unpackbits(b)[: c["m"]].sum()) for b in blooms] == c["bloom_popcounts"]
    built = O.OracleIndex.build(c["k"], c["m"], c["h"], blooms, c["samples"])
    assert np.array_equal(built.rows, ix.rows)


def test_synth_rows_shard_consistency():
    """Column shards of the synthetic matrix are slices of the one global matrix."""
    spec = O.SynthSpec(seed=7, and_draws=2, planted_cols=[3, 100, 1029, 4000], planted_thr=[0xFFFFFFFF, 1 << 31, 0xFFFFFFFF, 1 << 30])
    rows = np.array([0, 1, 5, 1000003, 24999999], dtype=np.int64)
    full = spec.rows(rows, 0, 4096)
    for off, n in ((0, 1024), (1024, 1024), (2048, 2048), (1000, 1000), (4088, 8)):
        part = spec.rows(rows, off, n)
        bits_full = np.unpackbits(full, axis=1)[:, off : off + n]
        bits_part = np.unpackbits(part, axis=1)[:, :n]
        assert np.array_equal(bits_full, bits_part)
        assert not np.unpackbits(part, axis=1)[:, n:].any()
    bits = np.unpackbits(full, axis=1)
    assert bits[:, 3].all() and bits[:, 1029].all()
    dens = np.unpackbits(spec.rows(np.arange(2000), 0, 4096), axis=1).mean()
    assert 0.22 < dens < 0.28


def test_counts_and_presence_consistency():
    rng = np.random.default_rng(0)
    m, N, h, k = 4096, 777, 3, 31
    rows = rng.integers(0, 256, size=(m, (N + 7) // 8), dtype=np.uint8)
    rows[:, -1] &= 0x80  # N % 8 == 1 -> 7 pad bits zero
    ix = O.OracleIndex(k, m, h, N, rows=rows)
    kmers = ["".join(rng.choice(list("ACGT"), size=k)) for _ in range(50)]
    packed = ix.lookup_packed(kmers)
    bits = np.unpackbits(packed, axis=1)[:, :N]
    assert np.array_equal(ix.counts(kmers), bits.sum(axis=0).astype(np.int32))
    assert np.array_equal(np.unpackbits(ix.presence(kmers))[:N], bits.all(axis=0).astype(np.uint8))
