"""GPU parity tests: the CUDA path, called through the C ABI (libbigsi_b200.so), against the CPU
oracle (oracle/) and the committed golden fixtures generated from the unmodified reference.
Everything here is bit-exact (integer / byte work)."""
import math
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden_util import bloom_from_b64, load, rows_from_b64

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import bigsi_b200

    assert bigsi_b200.device_count() >= 1, "no CUDA device visible to libbigsi_b200.so"
    return bigsi_b200


def _rand_kmers(rng, n, k, alphabet="ACGT"):
    a = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    return a[rng.integers(0, len(a), size=(n, k))]


def _kmer_strs(arr):
    return [bytes(r).decode() for r in arr]


# ---------------------------------------------------------------------------
# hashing (K1)
# ---------------------------------------------------------------------------
def test_hash_kat_and_golden(B):
    # /root/reference/bigsi/tests/bloom/test_create_bloomfilter.py:5-8
    assert B.generate_hashes("ATT", 3, 25) == {2, 15, 17}
    assert B.generate_hashes("ATT", 1, 25) == {15}
    assert B.generate_hashes("ATT", 2, 50) == {15, 27}
    g = load("hashes.json")
    for kat in g["kat"]:
        assert sorted(B.generate_hashes(kat["element"], kat["h"], kat["m"])) == kat["set"]
    for c in g["cases"]:
        got = B.hash_kmers([c["kmer"]], len(c["kmer"]), c["h"], c["m"])[0].tolist()
        assert got == c["rows"], c


@pytest.mark.parametrize("k", [1, 3, 4, 5, 31, 32, 33, 63])
def test_hash_random_vs_oracle(B, k):
    rng = np.random.default_rng(k)
    for alphabet in ("ACGT", "ACGTNacgt-"):
        arr = _rand_kmers(rng, 3000, k, alphabet)
        # add reverse-complement palindromes and pairs
        arr[1] = np.frombuffer(O.canonical(bytes(arr[0])), dtype=np.uint8)
        for h, m in ((1, 25), (3, 25_000_000), (5, 2_147_483_647), (2, 1), (4, 2), (3, 1000), (2, 1_073_741_827)):
            assert np.array_equal(B.hash_kmers(arr, k, h, m), O.hash_kmers(arr, k, h, m))
            nc = B.hash_kmers(arr, k, h, m, canonical=False)
            ref = np.array([[O.lib().oracle_hash_row(bytes(r), k, s, m) for s in range(h)] for r in arr[:50]])
            assert np.array_equal(nc[:50], ref)


# ---------------------------------------------------------------------------
# fused kernel vs oracle on random matrices
# ---------------------------------------------------------------------------
def _random_index(B, rng, m, N, density=0.5):
    rb = (N + 7) // 8
    rows = (rng.random((m, rb * 8)) < density)
    rows[:, N:] = False
    packed = np.packbits(rows, axis=1)
    ix = B.DeviceIndex(m, N)
    ix.upload_rows(0, packed)
    return ix, packed


def _check_batch(ix, packed, N, row_ids, qoff, h):
    counts = ix.search_rows(row_ids, h, q_offsets=qoff, mode=0)
    pres = ix.search_rows(row_ids, h, q_offsets=qoff, mode=1)
    for q in range(len(qoff) - 1):
        r = row_ids[qoff[q] : qoff[q + 1]]
        exp_c = O.counts_from_rows(packed, r, N)
        assert np.array_equal(counts[q].astype(np.int64), exp_c.astype(np.int64)), "counts q=%d" % q
        if len(r):
            exp_p = O.presence_from_rows(packed, r)
        else:
            exp_p = np.packbits(np.arange(((N + 7) // 8) * 8) < N)
        assert np.array_equal(pres[q], exp_p), "presence q=%d" % q


@pytest.mark.parametrize("N", [1, 7, 8, 64, 100, 127, 128, 129, 777, 4096, 5000])
@pytest.mark.parametrize("h", [1, 3])
def test_counts_and_presence_small(B, N, h):
    rng = np.random.default_rng(N * 10 + h)
    m = 5003
    ix, packed = _random_index(B, rng, m, N, density=0.7)
    for U in (1, 2, 7, 8, 9, 64, 1000):
        row_ids = rng.integers(0, m, size=(U, h), dtype=np.int32)
        _check_batch(ix, packed, N, row_ids, [0, U], h)
    ix.close()


@pytest.mark.parametrize("h", [2, 4, 5, 11, 40])
def test_other_hash_counts(B, h):
    rng = np.random.default_rng(h)
    m, N = 2000, 1500
    ix, packed = _random_index(B, rng, m, N, density=0.9)
    row_ids = rng.integers(0, m, size=(333, h), dtype=np.int32)
    row_ids[5] = row_ids[5, 0]  # duplicate hashes inside one k-mer
    _check_batch(ix, packed, N, row_ids, [0, 333], h)
    ix.close()


def test_multi_tile_wide_rows(B):
    """N > 53 248 columns -> more than one column tile per row."""
    rng = np.random.default_rng(5)
    m, N = 300, 120_001
    ix, packed = _random_index(B, rng, m, N, density=0.8)
    row_ids = rng.integers(0, m, size=(500, 3), dtype=np.int32)
    _check_batch(ix, packed, N, row_ids, [0, 500], 3)
    assert ix.info()["last_n_tiles"] >= 3
    ix.close()


def test_ragged_batch_with_empty_queries(B):
    rng = np.random.default_rng(11)
    m, N = 4000, 3001
    ix, packed = _random_index(B, rng, m, N, density=0.6)
    lens = [0, 1, 0, 0, 5, 300, 8, 0, 1024, 17, 0]
    qoff = np.concatenate([[0], np.cumsum(lens)])
    row_ids = rng.integers(0, m, size=(int(qoff[-1]), 3), dtype=np.int32)
    _check_batch(ix, packed, N, row_ids, qoff, 3)
    # many short queries (one merge slot each)
    lens = rng.integers(0, 40, size=500)
    qoff = np.concatenate([[0], np.cumsum(lens)])
    row_ids = rng.integers(0, m, size=(int(qoff[-1]), 3), dtype=np.int32)
    _check_batch(ix, packed, N, row_ids, qoff, 3)
    ix.close()


@pytest.mark.parametrize("h", [3, 1, 5])
def test_batch_shared_row_gather_reuse(B, h):
    """Batches whose queries share k-mers (BASELINE configs[4] "shared row-gather reuse"): the batch is de-duplicated by
    row-id tuple, the distinct tuples' AND vectors are gathered once and the queries count over those.  Counts, presence
    (AND mode) and thresholded hits against the oracle, identical to the path without reuse, through the k-mer and the
    row-id entry points; a batch of all-distinct k-mers does not take the reuse path."""
    rng = np.random.default_rng(211 + h)
    m, N, k = 7001, 1201, 31
    ix, packed = _random_index(B, rng, m, N, density=0.85)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    pool = _rand_kmers(rng, 2500, k)
    pool[1] = np.frombuffer(B.reverse_comp(bytes(pool[0]).decode()).encode(), dtype=np.uint8)  # same canonical k-mer, other string
    lens = [int(x) for x in rng.integers(300, 700, size=40)]
    lens[3] = 0
    qoff = np.concatenate([[0], np.cumsum(lens)])
    pick = np.concatenate([rng.choice(2500, size=n, replace=False) for n in lens if n] + [np.zeros(0, dtype=np.int64)])
    arr = np.ascontiguousarray(pool[pick.astype(np.int64)])
    assert arr.shape[0] == qoff[-1] >= 16384
    kmers = _kmer_strs(arr)
    mins = np.array([int(np.ceil(n * 0.6)) for n in lens], dtype=np.uint32)
    got = {}
    for reuse in (1, 0):
        ix.set_option("batch_reuse", reuse)
        counts = ix.search_kmers(arr, k, h, q_offsets=qoff)
        uniq = ix.info()["last_unique_kmers"]
        assert (0 < uniq <= 2500) if reuse else uniq == 0, uniq
        pres = ix.search_kmers(arr, k, h, q_offsets=qoff, mode=1)
        hits = ix.search_kmers_hits(arr, k, h, mins, q_offsets=qoff, cap=N)
        rows = O.hash_kmers(arr, k, h, m)
        counts_r = ix.search_rows(rows, h, q_offsets=qoff, mode=0)
        got[reuse] = (counts, pres)
        assert np.array_equal(counts, counts_r)
        for q, n in enumerate(lens):
            part = kmers[qoff[q]: qoff[q + 1]]
            cnt = oix.counts(part) if n else np.zeros(N, dtype=np.int64)
            assert np.array_equal(counts[q].astype(np.int64), cnt.astype(np.int64)), (reuse, q)
            if n:
                assert np.array_equal(pres[q], oix.presence(part)), (reuse, q)
            exp = np.nonzero(cnt >= mins[q])[0]
            cols, vals, nh = hits[q]
            assert nh == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals.astype(np.int64), cnt[exp]), (reuse, q)
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    ix.set_option("batch_reuse", 1)
    distinct = _rand_kmers(rng, 17_000, k)
    ix.search_kmers(distinct, k, h, q_offsets=[0, 9000, 17_000])
    if h >= 3:  # nothing to share: the ordinary batch path (with h = 1 the 7 001 rows ARE shared between 17 000 k-mers)
        assert ix.info()["last_unique_kmers"] == 0
    else:
        assert 0 < ix.info()["last_unique_kmers"] <= m
    ix.close()


@pytest.mark.parametrize("opts", [{}, {"grid": 7}, {"tile_bytes": 64, "grid": 5}, {"grid": 1}, {"fuse_merge": 0, "grid": 9}])
def test_batch_whole_query_segments_finished_directly(B, opts):
    """Batches: a query that lies inside one slice is finished by the CTA that counted it (no partial planes, no merge);
    queries cut by a slice boundary are merged.  Counts, thresholded hits (thresholds 0, mid, len, beyond len; a hit
    capacity smaller than the hit list) against the oracle, and identical to the all-merged path (option direct = 0).
    N is not a multiple of 8: the padding bits of the last byte must never become hits at threshold 0."""
    rng = np.random.default_rng(101)
    m, N, k, h, cap = 6007, 1003, 31, 3, 300
    ix, packed = _random_index(B, rng, m, N, density=0.8)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    lens = [40, 0, 1, 333, 7, 64, 0, 250, 19, 8, 100, 55, 2, 129, 31]
    qoff = np.concatenate([[0], np.cumsum(lens)])
    arr = _rand_kmers(rng, int(qoff[-1]), k)
    kmers = _kmer_strs(arr)
    mins = np.array([int(np.ceil(n * f)) for n, f in zip(lens, [0.5, 0.0, 1.0, 0.6, 0.0, 0.4, 1.0, 0.55, 1.0, 2.0, 0.5, 0.0, 0.5, 0.45, 0.5])],
                    dtype=np.uint32)
    results = {}
    for direct in (1, 0):
        ix.set_option("direct", direct)
        for key, v in opts.items():
            ix.set_option(key, v)
        counts = ix.search_kmers(arr, k, h, q_offsets=qoff)
        hits = ix.search_kmers_hits(arr, k, h, mins, q_offsets=qoff, cap=cap)
        results[direct] = (counts, hits)
        for q, n in enumerate(lens):
            cnt = oix.counts(kmers[qoff[q]: qoff[q + 1]]) if n else np.zeros(N, dtype=np.int64)
            assert np.array_equal(counts[q].astype(np.int64), cnt.astype(np.int64)), (direct, q)
            exp = np.nonzero(cnt >= mins[q])[0]
            cols, vals, nh = hits[q]
            assert nh == len(exp), (direct, q, nh, len(exp))
            if len(exp) <= cap:
                assert np.array_equal(cols, exp) and np.array_equal(vals.astype(np.int64), cnt[exp]), (direct, q)
            else:  # more hits than the list holds: any `cap` of them, each with its exact count
                assert len(cols) == cap and len(set(cols.tolist())) == cap
                assert all(cnt[c] == v and cnt[c] >= mins[q] for c, v in zip(cols.tolist(), vals.tolist())), (direct, q)
    assert np.array_equal(results[0][0], results[1][0])
    ix.close()


@pytest.mark.parametrize("opts", [
    {"tile_bytes": 16}, {"tile_bytes": 48, "kmers_per_stage": 1}, {"tile_bytes": 512, "kmers_per_stage": 4, "n_stages": 2},
    {"grid": 1}, {"grid": 3, "tile_bytes": 128}, {"grid": 1000}, {"kmers_per_stage": 8, "n_stages": 3},
    {"fuse_merge": 0}, {"fuse_merge": 0, "tile_bytes": 64}, {"fuse_merge": 0, "grid": 5}, {"grid": 148, "tile_bytes": 32},
])
def test_launch_geometry_overrides(B, opts):
    rng = np.random.default_rng(3)
    m, N = 3000, 2500
    ix, packed = _random_index(B, rng, m, N, density=0.75)
    for key, v in opts.items():
        ix.set_option(key, v)
    lens = [700, 0, 33, 1500]
    qoff = np.concatenate([[0], np.cumsum(lens)])
    row_ids = rng.integers(0, m, size=(int(qoff[-1]), 3), dtype=np.int32)
    _check_batch(ix, packed, N, row_ids, qoff, 3)
    ix.close()


def test_long_query_multi_slice(B):
    """More items than 65 535 per CTA: several slices per CTA and > 16 count planes."""
    rng = np.random.default_rng(17)
    m, N = 512, 200
    ix, packed = _random_index(B, rng, m, N, density=0.9)
    U = 148 * 65535 + 12345
    row_ids = rng.integers(0, m, size=(U, 3), dtype=np.int32)
    counts = ix.search_rows(row_ids, 3, mode=0)[0]
    assert np.array_equal(counts.astype(np.int64), O.counts_from_rows(packed, row_ids, N).astype(np.int64))
    assert counts.max() > 65535
    info = ix.info()
    assert info["last_n_slices"] > info["last_grid"]
    pres = ix.search_rows(row_ids, 3, mode=1)[0]
    assert np.array_equal(pres, O.presence_from_rows(packed, row_ids))
    ix.close()


def test_kmer_level_entry_points(B):
    rng = np.random.default_rng(23)
    m, N, k, h = 10_007, 1234, 31, 3
    ix, packed = _random_index(B, rng, m, N, density=0.8)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    arr = _rand_kmers(rng, 400, k)
    arr[7] = arr[3]  # a duplicate raw k-mer counts twice at this level (dedup is the caller's job)
    kmers = _kmer_strs(arr)
    assert np.array_equal(ix.search_kmers(arr, k, h)[0].astype(np.int64), oix.counts(kmers).astype(np.int64))
    assert np.array_equal(ix.search_kmers(arr, k, h, mode=1)[0], oix.presence(kmers))
    assert np.array_equal(ix.lookup_kmers(arr, k, h), oix.lookup_packed(kmers))
    # fused threshold
    cnt = oix.counts(kmers)
    for thr in (1, 100, 200, 400, 401):
        cols, vals, n = ix.search_kmers_hits(arr, k, h, [thr])[0]
        exp = np.nonzero(cnt >= thr)[0]
        assert n == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp])
    # capacity smaller than the number of hits: count is still exact
    cols, vals, n = ix.search_kmers_hits(arr, k, h, [1], cap=5)[0]
    assert n == int((cnt >= 1).sum()) and len(cols) == 5 and all(cnt[c] == v for c, v in zip(cols, vals))
    # batch of two queries with different thresholds
    res = ix.search_kmers_hits(arr, k, h, [150, 390], q_offsets=[0, 200, 400])
    for (cols, vals, n), sl, thr in zip(res, (slice(0, 200), slice(200, 400)), (150, 390)):
        c = oix.counts(kmers[sl])
        exp = np.nonzero(c >= thr)[0]
        assert np.array_equal(cols, exp) and np.array_equal(vals, c[exp])
    ix.close()


def test_set_column_and_download(B):
    rng = np.random.default_rng(29)
    m, N = 999, 13
    ix, packed = _random_index(B, rng, m, N, density=0.5)
    assert np.array_equal(ix.download_rows(0, m), packed)
    bits = np.unpackbits(packed, axis=1)[:, :N].copy()
    # overwrite column 4, then append three columns (one crosses a byte boundary: 13 -> 16)
    for col in (4, 13, 14, 15):
        bloom = rng.random(m) < 0.4
        ix.set_column(col, np.packbits(bloom), m)
        if col < bits.shape[1]:
            bits[:, col] = bloom
        else:
            bits = np.concatenate([bits, bloom[:, None]], axis=1)
    assert ix.num_cols == 16
    assert np.array_equal(np.unpackbits(ix.download_rows(0, m), axis=1)[:, :16], bits.astype(np.uint8))
    ix.close()


# ---------------------------------------------------------------------------
# synthetic index generator (K7) vs the oracle's pure function
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("col_offset,N", [(0, 1000), (50_000, 4096), (8, 77), (1_000_000 - 8, 50_000)])
def test_fill_synthetic_matches_oracle(B, col_offset, N):
    m = 700
    planted = [col_offset + 3, col_offset + N - 1, col_offset + N // 2, col_offset + N + 5, 1]
    thr = [0xFFFFFFFF, 1 << 31, 1 << 30, 0xFFFFFFFF, 0xFFFFFFFF]
    for and_draws in (1, 2):
        spec = O.SynthSpec(seed=42, and_draws=and_draws, planted_cols=planted, planted_thr=thr)
        ix = B.DeviceIndex(m, N, col_offset=col_offset)
        ix.fill_synthetic(42, and_draws, planted, thr)
        got = ix.download_rows(0, m)
        exp = spec.rows(np.arange(m), col_offset, N)
        assert np.array_equal(got, exp)
        ix.close()


# ---------------------------------------------------------------------------
# reference API: golden vectors from the unmodified reference
# ---------------------------------------------------------------------------
def _config(case, name, devices=None):
    sc = {"filename": name, "device": 0} if devices is None else {"filename": name, "devices": list(devices)}
    return {"k": case["k"], "m": case["m"], "h": case["h"], "storage-engine": "b200", "storage-config": sc}


# BIGSI(config) over one GPU, and column-sharded (storage-config.devices, bigsi_b200/sharded_index.py) over two and three
# shards -- on distinct GPUs where the box has them, else on the same one (the shard logic is the same)
def _device_sets():
    try:
        import torch

        n = torch.cuda.device_count()
    except Exception:
        n = 0
    return [None, [0, 1 % max(n, 1)], [0, 1 % max(n, 1), 2 % max(n, 1)]]


DEVICE_SETS = _device_sets()


def _check_queries(bigsi, queries):
    for q in queries:
        if "raises" in q:
            with pytest.raises(BaseException) as ei:
                bigsi.search(q["seq"], q["threshold"])
            assert type(ei.value).__name__ == q["raises"]
        else:
            assert bigsi.search(q["seq"], q["threshold"]) == q["result"], (q["seq"][:40], q["threshold"])


@pytest.mark.parametrize("devices", DEVICE_SETS, ids=lambda d: "1gpu" if d is None else "shards" + "".join(map(str, d)))
def test_reference_golden_search_cases(B, devices):
    g = load("search_cases.json")
    for ci, case in enumerate(g["cases"]):
        k, m, h = case["k"], case["m"], case["h"]
        cfg = _config(case, "golden-%d" % ci, devices)
        blooms = []
        for seq, ref_b64 in zip(case["sample_seqs"], case["blooms_b64"]):
            b = B.BIGSI.bloom(cfg, B.seq_to_kmers(seq, k))
            assert np.array_equal(np.frombuffer(b.tobytes(), dtype=np.uint8), bloom_from_b64(ref_b64))
            blooms.append(b)
        bigsi = B.BIGSI.build(cfg, blooms, case["samples"])
        assert np.array_equal(bigsi.index.download_rows(0, m), rows_from_b64(case["rows_b64"], m, case["row_bytes"]))
        _check_queries(bigsi, case["queries"])
        got = bigsi.lookup(case["lookup_in"])
        assert {km: v.to01() for km, v in got.items()} == case["lookup"]
        if "insert" in case:
            ins = case["insert"]
            bigsi.insert(B.BIGSI.bloom(cfg, B.seq_to_kmers(ins["seq"], k)), ins["sample"])
            assert bigsi.num_samples == ins["num_samples"]
            assert np.array_equal(bigsi.index.download_rows(0, m), rows_from_b64(ins["rows_b64"], m, ins["row_bytes"]))
            _check_queries(B.BIGSI(cfg), ins["queries"])
        if "delete" in case:
            d = case["delete"]
            bigsi.delete_sample(d["sample"])
            _check_queries(bigsi, d["queries"])
        bigsi.delete()
        with pytest.raises(BaseException):
            B.BIGSI(cfg)


def test_reference_end_to_end_kat(B):
    # /root/reference/bigsi/tests/graph/test_end_to_end.py:12-131 restated against this engine
    cfg = {"k": 3, "m": 1000, "h": 3, "storage-engine": "b200", "storage-config": {"filename": "kat"}}
    kat = load("search_cases.json")["reference_kat"]

    def mk(seqs, names):
        return B.BIGSI.build(cfg, [B.BIGSI.bloom(cfg, B.seq_to_kmers(s, 3)) for s in seqs], names)

    bigsi = mk(["ATACACAAT", "ACAGAGAAC"], ["a", "b"])
    assert bigsi.search("ATACACAAT")[0] == {"percent_kmers_found": 100, "num_kmers": 6, "num_kmers_found": 6, "sample_name": "a"}
    assert bigsi.search("ACAGAGAAC")[0] == {"percent_kmers_found": 100, "num_kmers": 6, "num_kmers_found": 6, "sample_name": "b"}
    assert bigsi.search("ACAGTTAAC") == []
    _check_queries(bigsi, kat["exact"])
    assert bigsi.kmer_size == 3 and bigsi.bloomfilter_size == 1000 and bigsi.num_hashes == 3 and bigsi.num_samples == 2
    with pytest.raises(ValueError):
        bigsi.insert(B.BIGSI.bloom(cfg, ["ATC"]), "a")  # duplicate sample name
    with pytest.raises(ValueError):
        B.BIGSI.build(cfg, [B.BIGSI.bloom(cfg, ["ATC"])], ["x", "y"])
    bigsi = mk(["ATACACAAT", "ATACACAAC"], ["a", "b"])
    _check_queries(bigsi, kat["inexact"])
    assert {k: v.to01() for k, v in bigsi.lookup("AAT").items()} == kat["inexact_lookup"]
    # test_index.py:14-44 lookups incl. canonicalisation and duplicate k-mers
    res = bigsi.lookup(["ATA", "ATA", "TAT"])
    assert res["ATA"] == res["TAT"] and len(res) == 2 and len(res["ATA"]) == 2
    assert len(bigsi.lookup("ATA", remove_trailing_zeros=False)["ATA"]) == 8
    with pytest.raises(AssertionError):
        bigsi.search("ATACACAAT", 1.5)
    bigsi.delete()


@pytest.mark.parametrize("devices", DEVICE_SETS, ids=lambda d: "1gpu" if d is None else "shards" + "".join(map(str, d)))
def test_config1_golden(B, devices):
    c = load("config1.json")
    cfg = _config(c, "config1", devices)
    blooms = [B.BIGSI.bloom(cfg, km) for km in c["sample_kmers"]]
    assert [b.count() for b in blooms] == c["bloom_popcounts"]
    bigsi = B.BIGSI.build(cfg, blooms, c["samples"])
    assert np.array_equal(bigsi.index.download_rows(0, c["m"]), rows_from_b64(c["rows_b64"], c["m"], c["row_bytes"]))
    _check_queries(bigsi, c["queries"])
    bigsi.delete()


@pytest.mark.parametrize("devices", DEVICE_SETS, ids=lambda d: "1gpu" if d is None else "shards" + "".join(map(str, d)))
def test_random_api_parity_vs_oracle(B, devices):
    """Randomised BIGSI.search vs the oracle restatement, incl. non-ACGT bases and reverse
    complements in one query, thresholds 0..1, N not a multiple of 8."""
    rng = np.random.default_rng(101)
    k, m, h, n = 11, 20_011, 3, 37
    cfg = _config({"k": k, "m": m, "h": h}, "rand-api", devices)
    genomes = ["".join(rng.choice(list("ACGT"), size=600)) for _ in range(n)]
    for i in range(1, n, 3):  # related samples: shared prefixes
        genomes[i] = genomes[i - 1][:400] + genomes[i][400:]
    names = ["s%d" % i for i in range(n)]
    blooms = [B.BIGSI.bloom(cfg, B.seq_to_kmers(g, k)) for g in genomes]
    bigsi = B.BIGSI.build(cfg, blooms, names)
    oblooms = [O.OracleIndex.bloom(k, m, h, [O.canonical(x) for x in O.seq_to_kmers(g, k)]) for g in genomes]
    oix = O.OracleIndex.build(k, m, h, oblooms, names)
    assert np.array_equal(bigsi.index.download_rows(0, m), oix.rows)
    queries = [genomes[0][:200], genomes[1][350:450], genomes[4][100:130] + "N" + genomes[4][131:180],
               B.reverse_comp(genomes[7][:90]) + genomes[7][:90], genomes[9][:k], "ACGT" * 20]
    for qseq in queries:
        for thr in (1.0, 0.9, 0.5, 0.31, 0.0):
            assert bigsi.search(qseq, thr) == oix.search(qseq, thr), (qseq[:20], thr)
    bigsi.delete()


# ---------------------------------------------------------------------------
# BASELINE config 2 at full size: m = 25 M, N = 50 000 (156.8 GB in HBM)
# ---------------------------------------------------------------------------
def test_full_size_config2_parity(B):
    import torch

    free, total = torch.cuda.mem_get_info(0)
    m, N, k, h = 25_000_000, 50_000, 31, 3
    if free < m * 6272 + (4 << 30):
        pytest.skip("needs %.0f GB of free HBM" % (m * 6272 / 1e9))
    planted = [0, 1, 49_999, 25_000, 7, 12_345, 33_333]
    thr = [0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, int(0.95 * 2 ** 32), int(0.6 * 2 ** 32), int(0.41 * 2 ** 32), int(0.3 * 2 ** 32)]
    ix = B.DeviceIndex(m, N)
    ix.fill_synthetic(0, 1, planted, thr)
    spec = O.SynthSpec(0, 1, planted, thr)
    oix = O.OracleIndex(k, m, h, N, synth=spec)
    rng = np.random.default_rng(1)
    arr = _rand_kmers(rng, 10_000, k)
    kmers = _kmer_strs(arr)
    # spot-check stored rows incl. the last one
    probe = np.array([0, 1, 12_345_678, m - 1])
    for r in probe:
        assert np.array_equal(ix.download_rows(int(r), 1)[0], spec.rows([r], 0, N)[0])
    exp = oix.counts(kmers)
    got = ix.search_kmers(arr, k, h)[0]
    assert np.array_equal(got.astype(np.int64), exp.astype(np.int64))
    assert got[0] == 10_000 and got[1] == 10_000 and got[49_999] == 10_000
    assert np.array_equal(ix.search_kmers(arr, k, h, mode=1)[0], oix.presence(kmers))
    # exact query == threshold at U; 0.4 threshold keeps the graded planted columns
    cols, vals, n = ix.search_kmers_hits(arr, k, h, [10_000])[0]
    assert cols.tolist() == [0, 1, 49_999]
    mk = math.ceil(10_000 * 0.4)
    cols, vals, n = ix.search_kmers_hits(arr, k, h, [mk])[0]
    e = np.nonzero(exp >= mk)[0]
    assert np.array_equal(cols, e) and np.array_equal(vals, exp[e]) and 25_000 in cols.tolist()
    # size-independent properties: counts are additive over a split of the query ...
    a = ix.search_kmers(arr[:3777], k, h)[0].astype(np.int64)
    b = ix.search_kmers(arr[3777:], k, h)[0].astype(np.int64)
    assert np.array_equal(a + b, got.astype(np.int64))
    # ... and batched == one by one
    both = ix.search_kmers(arr, k, h, q_offsets=[0, 3777, 10_000])
    assert np.array_equal(both[0], a) and np.array_equal(both[1], b)
    ix.close()


# ---------------------------------------------------------------------------
# device-pointer entry points on torch tensors (the multi-GPU plumbing path)
# ---------------------------------------------------------------------------
def test_device_pointer_path_fused_threshold(B):
    import torch

    from bigsi_b200.sharded import DeviceShard, ShardedSearcher, unpack_hits

    rng = np.random.default_rng(31)
    m, N, k, h = 30_011, 9000, 31, 3
    ix, packed = _random_index(B, rng, m, N, density=0.85)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    shard = DeviceShard(ix, k, h, cap=512)
    lens = [300, 0, 1, 900]
    qoff = np.concatenate([[0], np.cumsum(lens)])
    arr = _rand_kmers(rng, int(qoff[-1]), k)
    kmers = _kmer_strs(arr)
    mins = np.array([150, 1, 1, 520], dtype=np.int32)
    dev = shard.device
    d_k = torch.from_numpy(arr).to(dev)
    d_q = torch.from_numpy(qoff.astype(np.int64)).to(dev)
    d_min = torch.from_numpy(mins).to(dev)
    rows = shard.hash(d_k)
    assert np.array_equal(rows.cpu().numpy(), O.hash_kmers(arr, k, h, m))
    # unfused: counts then threshold
    counts = shard.counts(rows, d_q, 4, max(lens))
    n1, c1, v1 = unpack_hits(shard.hits(counts, d_min).cpu().numpy(), 4, 512)
    # fused
    n2, c2, v2 = unpack_hits(shard.search_hits(rows, d_q, 4, d_min, max(lens)).cpu().numpy(), 4, 512)
    g = ShardedSearcher(shard).search_step(d_k, d_q, d_min, 4, max(lens))
    n3, c3, v3 = unpack_hits(g.cpu().numpy(), 4, 512)
    assert ix.info()["last_fused"] & 1  # merge ran inside the fused kernel
    # the same through the separate hash / merge kernels
    ix.set_option("fuse_merge", 0)
    ix.set_option("prehash", 0)
    n4, c4, v4 = unpack_hits(shard.search_kmers_hits(d_k, d_q, 4, d_min, max(lens)).cpu().numpy(), 4, 512)
    assert ix.info()["last_fused"] == 0
    ix.set_option("fuse_merge", 1)
    ix.set_option("prehash", 1)
    assert np.array_equal(n3, n4)
    torch.cuda.synchronize()
    for q in range(4):
        sl = slice(qoff[q], qoff[q + 1])
        cnt = oix.counts(kmers[sl]) if lens[q] else np.zeros(N, dtype=np.int32)
        assert np.array_equal(counts[q, :N].cpu().numpy().astype(np.int64), cnt.astype(np.int64))
        exp = np.nonzero(cnt >= mins[q])[0]
        for n, c, v in ((n1, c1, v1), (n2, c2, v2), (n3, c3, v3)):
            assert n[0, q] == len(exp)
            got = min(len(exp), 512)
            order = np.argsort(c[0, q, :got])
            if len(exp) <= 512:
                assert np.array_equal(c[0, q, :got][order], exp) and np.array_equal(v[0, q, :got][order], cnt[exp])
            else:
                assert all(cnt[cc] == vv and cnt[cc] >= mins[q] for cc, vv in zip(c[0, q, :got], v[0, q, :got]))
    ix.close()


# ---------------------------------------------------------------------------
# single-query host path: zero-copy k-mers, threshold by value, kernel-published result block
# ---------------------------------------------------------------------------
def _pinned_copy(B, arr):
    """arr copied into a buffer from bigsi_b200_host_alloc (mapped pinned memory); returns (view, free)."""
    import ctypes

    from bigsi_b200 import _lib

    L = _lib.lib()
    p = ctypes.c_void_p(0)
    _lib.check(L.bigsi_b200_host_alloc(max(arr.nbytes, 1), ctypes.byref(p)))
    buf = (ctypes.c_uint8 * max(arr.nbytes, 1)).from_address(p.value)
    view = np.frombuffer(buf, dtype=np.uint8, count=arr.nbytes).reshape(arr.shape)
    view[...] = arr
    return view, lambda: L.bigsi_b200_host_free(p)


@pytest.mark.parametrize("pinned", [False, True])
def test_single_query_zero_copy_path(B, pinned):
    rng = np.random.default_rng(41)
    m, N, k, h = 10_007, 5000, 31, 3
    ix, packed = _random_index(B, rng, m, N, density=0.9)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    free = None
    try:
        for n_kmers in (1, 37, 1500, 6000):
            arr = _rand_kmers(rng, n_kmers, k)
            if pinned:
                arr, free_now = _pinned_copy(B, arr)
            cnt = oix.counts(_kmer_strs(arr))
            for frac in (1.0, 0.7, 0.0):
                thr = int(math.ceil(n_kmers * frac))
                exp = np.nonzero(cnt >= thr)[0]
                # the default capacity (N) exceeds the 1024 hits the published block holds: long lists
                # come back through the device buffers
                cols, vals, n = ix.search_kmers_hits(arr, k, h, [thr])[0]
                assert n == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp]), (n_kmers, frac)
                cols, vals, n = ix.search_kmers_hits(arr, k, h, [thr], cap=64)[0]
                assert n == len(exp) and len(cols) == min(64, len(exp))
                assert all(cnt[c] == v and v >= thr for c, v in zip(cols, vals))
            assert ix.info()["last_fused"] & 8  # streamed launch: gather kernel + flush kernel
            # same answers from the staged path
            ix.set_option("zero_copy", 0)
            thr = int(math.ceil(n_kmers * 0.7))
            exp = np.nonzero(cnt >= thr)[0]
            cols, vals, n = ix.search_kmers_hits(arr, k, h, [thr])[0]
            assert n == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp])
            ix.set_option("zero_copy", 1)
            if pinned:
                free_now()
    finally:
        ix.close()


@pytest.mark.parametrize("opts", [
    {"pool_pct": 0}, {"pool_pct": 50}, {"pool_pct": 100}, {"solo": 0}, {"n_stages": 2}, {"grid": 3},
    {"merge_chunk_bytes": 16}, {"merge_chunk_bytes": 1024}, {"kmers_per_stage": 4, "pool_pct": 30},
    {"self_merge": 1}, {"self_merge": 1, "pool_pct": 0, "n_stages": 3}, {"defer": 0},
])
def test_solo_path_geometries(B, opts):
    """The single-query in-kernel path (producer-warp hashing, pooled tail k-mers, chunk-major merge)
    under forced geometries; counts and hits must not depend on any of them."""
    rng = np.random.default_rng(43)
    m, N, k, h = 10_007, 7000, 31, 3
    ix, packed = _random_index(B, rng, m, N, density=0.9)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    for key, val in opts.items():
        ix.set_option(key, val)
    try:
        for n_kmers in (5, 148, 3000, 9000):
            arr = _rand_kmers(rng, n_kmers, k)
            cnt = oix.counts(_kmer_strs(arr))
            thr = int(math.ceil(n_kmers * 0.75))
            exp = np.nonzero(cnt >= thr)[0]
            cols, vals, n = ix.search_kmers_hits(arr, k, h, [thr])[0]
            assert n == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp]), (opts, n_kmers)
            assert np.array_equal(ix.search_kmers(arr, k, h)[0].astype(np.int64), cnt.astype(np.int64))
            assert np.array_equal(ix.search_kmers(arr, k, h, mode=1)[0], oix.presence(_kmer_strs(arr)))
    finally:
        ix.close()


def test_deferred_stream_of_single_queries(B):
    """The deferred entry point (bigsi_b200_query_kmers_hits_stream_dev + bigsi_b200_index_flush): stage 2 of query s is
    executed by the merge team of query s+1's kernel, or by the flush kernel.  A burst of queries of very different
    sizes without any host synchronisation -- tiny grids (several kernels resident at once), queries too long for the
    team variant (flushed by the next launch), an AND-mode call and a batch in between (both flush first), ring slots
    and state blocks rotating several times -- every hit list against the oracle, and identical with option defer = 0."""
    import torch

    from bigsi_b200.sharded import DeviceShard, unpack_hits

    rng = np.random.default_rng(307)
    m, N, k, h, cap = 20_011, 6000, 31, 3, 4096
    ix, packed = _random_index(B, rng, m, N, density=0.9)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    shard = DeviceShard(ix, k, h, cap=cap)
    dev = shard.device
    sizes = [3000, 1, 40, 9000, 148, 60_000, 5, 7000, 2500, 149, 33, 45_000, 4000, 4000, 1, 6000, 800, 12_000, 7, 3000, 3000, 90]
    queries = [_rand_kmers(rng, n, k) for n in sizes]
    d_queries = [torch.from_numpy(a).to(dev) for a in queries]
    expected = []
    for a, n in zip(queries, sizes):
        cnt = oix.counts(_kmer_strs(a))
        thr = int(math.ceil(n * 0.8))
        expected.append((thr, cnt, np.nonzero(cnt >= thr)[0]))
    d_and = torch.empty((1, (N + 7) // 8 + 16), dtype=torch.uint8, device=dev)
    d_qoff1 = torch.tensor([0, sizes[0]], dtype=torch.int64, device=dev)
    ix.set_option("inputs_ready", 1)
    try:
        for defer in (1, 0):
            ix.set_option("defer", defer)
            bufs = []
            for j, n in enumerate(sizes):
                bufs.append(shard.search_kmers_hits_stream(d_queries[j], expected[j][0]))
                if j % 8 == 7:
                    # the ring of output buffers has 8 entries: consume before they come round again.  The NEWEST result
                    # is only complete after the next streamed query or a flush; the seven before it are complete now
                    # (in stream order), which is what this copy relies on -- then flush for the eighth.
                    done = [b.clone() for b in bufs[-8:-1]]
                    shard.flush()
                    done.append(bufs[-1].clone())
                    bufs[-8:] = done
                if j == 4:  # an AND-mode query on the same handle: flushes the pending query first
                    rows = shard.hash(d_queries[0])
                    ix.query_dev(1, rows.data_ptr(), d_qoff1.data_ptr(), 1, sizes[0], h, d_and.data_ptr(), d_and.shape[1],
                                 torch.cuda.current_stream().cuda_stream, sizes[0])
                if j == 10:  # a batch (generic kernel) in between
                    both = torch.cat([d_queries[2], d_queries[4]])
                    qo = torch.tensor([0, sizes[2], sizes[2] + sizes[4]], dtype=torch.int64, device=dev)
                    mn = torch.tensor([expected[2][0], expected[4][0]], dtype=torch.int32, device=dev)
                    batch = shard.search_kmers_hits(both, qo, 2, mn, max(sizes[2], sizes[4])).clone()
            shard.flush()
            tail = [b.clone() for b in bufs[len(sizes) - len(sizes) % 8:]]
            bufs[len(sizes) - len(sizes) % 8:] = tail
            torch.cuda.synchronize()
            assert B._lib.lib().bigsi_b200_index_status(ix.handle) == 0
            for j, n in enumerate(sizes):
                thr, cnt, exp = expected[j]
                nh, cols, vals = unpack_hits(bufs[j].cpu().numpy(), 1, cap)
                assert int(nh[0, 0]) == len(exp), (defer, j, n, int(nh[0, 0]), len(exp))
                got = min(int(nh[0, 0]), cap)
                order = np.argsort(cols[0, 0, :got])
                if len(exp) <= cap:
                    assert np.array_equal(cols[0, 0, :got][order], exp) and np.array_equal(vals[0, 0, :got][order], cnt[exp]), (defer, j)
                else:  # more hits than the list holds (a 1-k-mer query): any `cap` distinct hits with their exact counts
                    c = cols[0, 0, :got]
                    assert len(set(c.tolist())) == cap and np.array_equal(vals[0, 0, :got], cnt[c]) and (cnt[c] >= thr).all(), (defer, j)
            assert np.array_equal(d_and[0, : (N + 7) // 8].cpu().numpy(), oix.presence(_kmer_strs(queries[0])))
            nb, cb, vb = unpack_hits(batch.cpu().numpy(), 2, cap)
            for q, j in enumerate((2, 4)):
                got = int(nb[0, q])
                assert got == len(expected[j][2]) and np.array_equal(np.sort(cb[0, q, :got]), expected[j][2]), (defer, "batch", q)
    finally:
        ix.close()


# ---------------------------------------------------------------------------
# column-sharded search: query broadcast + hit all-gather fused into the query kernels
# ---------------------------------------------------------------------------
def _make_shards(B, world, m, part, k, h, cap, packed, max_kmers, opts=None):
    """`world` column shards of `part` columns each, one handle per shard, wired through plain peer access
    inside this process.  Shard g lives on device g % device_count: on a box with fewer GPUs than shards several
    handles share a device (each with its own stream), which exercises the same protocol."""
    import torch

    from bigsi_b200.sharded import DeviceShard, FusedExchange

    ndev = torch.cuda.device_count()
    shards, exs, streams = [], [], []
    for g in range(world):
        ix = B.DeviceIndex(m, part, col_offset=g * part, device=g % ndev)
        ix.upload_rows(0, packed, src_byte_offset=g * part // 8)
        for key, val in (opts or {}).items():
            ix.set_option(key, val)
        shards.append(DeviceShard(ix, k, h, cap=cap))
        streams.append(torch.cuda.Stream(device=g % ndev))
    for g in range(world):
        exs.append(FusedExchange(shards[g], world, g, max_kmers, peers=True))
    FusedExchange.connect_local(exs)
    return shards, exs, streams


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_fused_exchange_shards_one_process(B, world):
    """Streamed exchange over `world` shards in one process: a burst of back-to-back queries without any host
    synchronisation in between (consecutive queries overlap on every shard, inboxes / generations / ring slots
    rotate several times), every shard's copy of every query's all-gathered hits against the oracle."""
    import torch

    from bigsi_b200.sharded import merge_shard_hits, unpack_hits

    ndev = torch.cuda.device_count()
    rng = np.random.default_rng(53 + world)
    part = 2000  # columns per shard, a multiple of 8
    m, N, k, h, cap = 20_011, part * world, 31, 3, 2048
    rb = (N + 7) // 8
    rows = rng.random((m, rb * 8)) < 0.9
    rows[:, N:] = False
    packed = np.packbits(rows, axis=1)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    offs = [g * part for g in range(world)]
    # shards that share a device: small rings (n_stages), so that the CTAs of all shards fit on the SMs together --
    # a shard whose CTAs wait (for rank 0's k-mers, at the entry gate) must not keep the others off the device --
    # and rank 0 first; one shard per device: the peers first, so that their kernels really wait for rank 0's push
    shared = ndev < world
    opts = {"inputs_ready": 1, "n_stages": 2} if shared else {"inputs_ready": 1}
    shards, exs, streams = _make_shards(B, world, m, part, k, h, cap, packed, 8000, opts)
    order = list(range(world)) if shared else list(range(world - 1, -1, -1))
    try:
        for g in range(world):  # first use of torch's copy kernel / allocator must not fall between the ranks' launches
            with torch.cuda.device(shards[g].device), torch.cuda.stream(streams[g]):
                torch.zeros((world, 2 + 2 * cap), dtype=torch.int32, device=shards[g].device).clone()
        for d in range(ndev):
            torch.cuda.synchronize(d)
        sizes = [40, 3000, 1, 7000, 512, 2500, 6000, 90, 4000, 4000, 333, 8000, 17, 5000]
        queries = [_rand_kmers(rng, n, k) for n in sizes]
        d_queries = [torch.from_numpy(a).to(shards[0].device) for a in queries]
        torch.cuda.synchronize(shards[0].device)
        for base, end in ((0, 1), (1, 7), (7, len(sizes))):  # 1, 6 and 7 queries in flight without a host sync
            burst = sizes[base:end]
            copies = []
            prev = None

            def consume(views):  # on each shard's own stream, in stream order
                per_rank = [None] * world
                for g in order:
                    with torch.cuda.device(shards[g].device), torch.cuda.stream(streams[g]):
                        per_rank[g] = views[g].clone()
                copies.append(per_rank)

            for j, n_kmers in enumerate(burst):
                thr = int(math.ceil(n_kmers * 0.85))
                views = [None] * world
                for g in order:  # every rank's launch first: nothing that may block the host in between
                    with torch.cuda.device(shards[g].device), torch.cuda.stream(streams[g]):
                        views[g] = exs[g].search(d_queries[base + j] if g == 0 else None, n_kmers, thr)
                if prev is not None:  # deferred: query j-1 is complete behind the launch of query j
                    consume(prev)
                prev = views
            for g in order:  # the last query of the burst: stage 2 as a kernel of its own, on every shard
                exs[g].flush()
            consume(prev)
            for d in range(ndev):
                torch.cuda.synchronize(d)
            for g in range(world):
                assert shards[g].index.info()["last_fused"] & 8
            for j, n_kmers in enumerate(burst):
                thr = int(math.ceil(n_kmers * 0.85))
                cnt = oix.counts(_kmer_strs(queries[base + j]))
                exp = np.nonzero(cnt >= thr)[0]
                for g in range(world):  # every shard holds every shard's hits (all-gather)
                    n, cols, vals = unpack_hits(copies[j][g].cpu().numpy(), 1, cap)
                    assert (n[:, 0] <= cap).all()
                    gc, gv = merge_shard_hits(n[:, 0], cols[:, 0], vals[:, 0], offs)
                    assert np.array_equal(gc, exp) and np.array_equal(gv, cnt[exp]), (world, base + j, g)
        w, q = exs[0].wait_ns()
        assert q == len(sizes)
    finally:
        for e in exs:
            e.close()
        for s in shards:
            s.index.close()


def test_exchange_times_out_instead_of_hanging(B):
    """A rank whose peer never launches its search gets an error code within a bounded time (no hung GPU):
    rank 0's stage 2 gives up waiting for rank 1's hit list, the handle reports BIGSI_B200_ERR_TIMEOUT
    from then on; a peer that waits for k-mers nobody sends times out the same way."""
    import time

    import torch

    from bigsi_b200 import _lib

    rng = np.random.default_rng(59)
    part, m, k, h, cap = 800, 5003, 31, 3, 256
    rows = rng.random((m, 2 * part)) < 0.9
    packed = np.packbits(rows, axis=1)
    for silent in (1, 0):
        shards, exs, streams = _make_shards(B, 2, m, part, k, h, cap, packed, 1000, {"spin_timeout_ms": 300, "inputs_ready": 1})
        try:
            arr = _rand_kmers(rng, 500, k)
            d_k = torch.from_numpy(arr).to(shards[0].device)
            talker = 1 - silent
            t0 = time.perf_counter()
            with torch.cuda.device(shards[talker].device), torch.cuda.stream(streams[talker]):
                exs[talker].search(d_k if talker == 0 else None, 500, 400)
                exs[talker].search(d_k if talker == 0 else None, 500, 400)  # a second query behind the stuck one
            streams[talker].synchronize()
            dt = time.perf_counter() - t0
            assert dt < 5.0, "the stuck search took %.1f s to give up" % dt
            rc = _lib.lib().bigsi_b200_index_status(shards[talker].index.handle)
            assert rc == _lib.ERR_TIMEOUT
            msg = _lib.lib().bigsi_b200_last_error().decode()
            assert "timed out" in msg
            with pytest.raises(_lib.BigsiB200Error) as ei:  # sticky: later calls fail instead of running
                with torch.cuda.device(shards[talker].device), torch.cuda.stream(streams[talker]):
                    exs[talker].search(d_k if talker == 0 else None, 500, 400)
            assert ei.value.code == _lib.ERR_TIMEOUT
            # the silent rank never launched anything: it is healthy
            assert _lib.lib().bigsi_b200_index_status(shards[silent].index.handle) == 0
        finally:
            for e in exs:
                e.close()
            for s in shards:
                s.index.close()


def test_exchange_across_processes_ipc(B, tmp_path):
    """The mode bench.py's multi-GPU runs use: one PROCESS per shard, result blocks and inboxes mapped through
    CUDA IPC (bigsi_b200_exchange_create / open).  world = min(GPUs, 8) processes (two processes sharing the GPU
    on a single-GPU box), >= 12 queries back to back, every rank's merged (colour, count) list against the oracle
    (tests/exchange_worker.py does the checking and writes one verdict per rank)."""
    import json
    import socket
    import subprocess
    import sys

    import torch

    ndev = torch.cuda.device_count()
    world = min(ndev, 8) if ndev >= 2 else 2
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r % ndev), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(os.path.dirname(__file__), "exchange_worker.py"),
                                       str(tmp_path)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for pr in procs:
        try:
            out, _ = pr.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, pr in enumerate(procs):
        assert pr.returncode == 0, "rank %d failed:\n%s" % (r, outs[r][-3000:])
    for r in range(world):
        with open(os.path.join(str(tmp_path), "rank%d.json" % r)) as f:
            verdict = json.load(f)
        assert verdict["ok"] and verdict["queries"] >= 12 and verdict["world"] == world, verdict


# ---------------------------------------------------------------------------
# BASELINE configs 3-5 at full per-GPU size (one 50 000-column shard of the ENA-scale index)
# ---------------------------------------------------------------------------
_PLANTED = [0, 1, 49_999, 25_000, 7, 12_345, 33_333]
_PLANTED_THR = [0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, int(0.95 * 2 ** 32), int(0.6 * 2 ** 32), int(0.41 * 2 ** 32),
                int(0.3 * 2 ** 32)]


@pytest.fixture(scope="module")
def big_shard(B):
    """Shard 3 of the 8-shard config-3 index: columns [150 000, 200 000) of N = 400 000, m = 25 M."""
    import torch

    free, total = torch.cuda.mem_get_info(0)
    m, N, off = 25_000_000, 50_000, 150_000
    if free < m * 6272 + (6 << 30):
        pytest.skip("needs %.0f GB of free HBM" % (m * 6272 / 1e9))
    planted = [off + c for c in _PLANTED]
    ix = B.DeviceIndex(m, N, col_offset=off)
    ix.fill_synthetic(0, 1, planted, _PLANTED_THR)
    oix = O.OracleIndex(31, m, 3, N, synth=O.SynthSpec(0, 1, planted, _PLANTED_THR), col_offset=off)
    yield ix, oix
    ix.close()


def _unique_kmers_of(seq_arr, k):
    """Raw unique k-mers of an ACGT sequence (set semantics, graph/index.py:45), first-occurrence order."""
    assert k <= 32
    win = np.lib.stride_tricks.sliding_window_view(seq_arr, k)
    code = np.zeros(256, dtype=np.uint64)
    code[[65, 67, 71, 84]] = [0, 1, 2, 3]
    packed = (code[win] << (2 * np.arange(k, dtype=np.uint64))).sum(axis=1, dtype=np.uint64)  # exact 2-bit key
    _, first = np.unique(packed, return_index=True)
    return np.ascontiguousarray(win[np.sort(first)])


def test_config3_megabase_query_full_size(B, big_shard):
    """1 Mbp query (999 970 k-mers) on one shard: additivity against verified 10 000-k-mer pieces,
    planted columns, AND mode, and the threshold path of config 4 on the same counts."""
    ix, oix = big_shard
    k, h = 31, 3
    rng = np.random.default_rng(2)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=1_000_000)]
    arr = _unique_kmers_of(seq, k)
    U = arr.shape[0]
    assert 999_000 < U <= 999_970
    got = ix.search_kmers(arr, k, h)[0].astype(np.int64)
    assert ix.info()["last_n_slices"] >= 148
    # all-ones planted columns count every k-mer; the graded ones are ordered by density
    assert got[0] == U and got[1] == U and got[49_999] == U
    assert got[25_000] > got[7] > got[100] > got[12_345] > got[33_333]  # densities .95 > .6 > .5 (plain) > .41 > .3
    # additivity: the sum over 100 disjoint pieces must reproduce the big query exactly ...
    piece = 10_000
    acc = np.zeros_like(got)
    verified = 0
    for i, p0 in enumerate(range(0, U, piece)):
        part = arr[p0 : p0 + piece]
        c = ix.search_kmers(part, k, h)[0].astype(np.int64)
        if i in (0, 57):  # ... and pieces are checked against the oracle (rows regenerated on the CPU)
            assert np.array_equal(c, oix.counts(_kmer_strs(part)).astype(np.int64))
            verified += 1
        acc += c
    assert verified == 2 and np.array_equal(acc, got)
    # exact filter == AND mode == threshold at U
    pres = np.unpackbits(ix.search_kmers(arr, k, h, mode=1)[0])[:50_000]
    assert np.array_equal(np.nonzero(pres)[0], np.nonzero(got == U)[0])
    cols, vals, n = ix.search_kmers_hits(arr, k, h, [U])[0]
    assert cols.tolist() == [0, 1, 49_999] and n == 3
    # config 4: score >= 0.4 on the megabase query
    mk = math.ceil(U * 0.4)
    cols, vals, n = ix.search_kmers_hits(arr, k, h, [mk])[0]
    e = np.nonzero(got >= mk)[0]
    assert np.array_equal(cols, e) and np.array_equal(vals.astype(np.int64), got[e])
    assert cols.tolist() == [0, 1, 25_000, 49_999]  # 0.95^3 = 0.86 >= 0.4; 0.6^3 = 0.22 and the rest are below


def test_config5_batch_of_1000_queries_full_size(B, big_shard):
    """1 000 queries x 1 000 k-mers in one launch, independent and heavily overlapping variants:
    batch == one by one == oracle; thresholds per query."""
    ix, oix = big_shard
    k, h, Q, L = 31, 3, 1000, 1000
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    indep = acgt[rng.integers(0, 4, size=(Q * L, k))]
    base = acgt[rng.integers(0, 4, size=100_000 + k)]
    win = np.lib.stride_tricks.sliding_window_view(base, k)
    starts = rng.integers(0, 100_000 - L, size=Q)
    shared = np.ascontiguousarray(np.concatenate([win[s : s + L] for s in starts]))  # windows of one sequence
    qoff = np.arange(0, Q * L + 1, L)
    for name, arr in (("independent", indep), ("shared", shared)):
        counts = ix.search_kmers(arr, k, h, q_offsets=qoff)
        assert counts.shape[0] == Q
        assert (counts[:, 0] == L).all() and (counts[:, 49_999] == L).all()
        for q in (0, 499, 999):
            part = arr[q * L : (q + 1) * L]
            assert np.array_equal(counts[q].astype(np.int64), oix.counts(_kmer_strs(part)).astype(np.int64)), (name, q)
            assert np.array_equal(ix.search_kmers(part, k, h)[0], counts[q]), (name, q)
        mins = np.full(Q, math.ceil(L * 0.4), dtype=np.uint32)
        mins[::2] = L
        res = ix.search_kmers_hits(arr, k, h, mins, q_offsets=qoff, cap=256)
        for q in (0, 1, 2, 777):
            cols, vals, n = res[q]
            e = np.nonzero(counts[q] >= mins[q])[0]
            assert n == len(e) and np.array_equal(cols, e) and np.array_equal(vals, counts[q][e]), (name, q)


# ---------------------------------------------------------------------------
# native query front-end: sequence -> set of raw k-mers -> search, one C-ABI call
# ---------------------------------------------------------------------------
def test_search_sequence_front_end(B):
    rng = np.random.default_rng(53)
    m, N, h = 10_007, 3000, 3
    ix, packed = _random_index(B, rng, m, N, density=0.9)
    try:
        for k in (5, 11, 31, 32):
            oix = O.OracleIndex(k, m, h, N, rows=packed)
            base = "".join(rng.choice(list("ACGT"), size=700))
            seqs = [
                base,                                # all windows distinct (k >= 11) or heavily repeated (k = 5)
                base[:300] + base[:300] + base[100:250],   # repeats: duplicate raw k-mers count once
                "ACGT" * 40,                         # four distinct windows at most
                base[:120].lower() + "N" + base[120:200] + "-" + base[:50],  # non-ACGT bytes pass through
                base[:k],                            # exactly one window
                "A" * (k + 30),                      # one unique k-mer, many windows
            ]
            for seq in seqs:
                uk = O.unique_kmers(seq, k)
                cnt = oix.counts(uk)
                for thr in (1.0, 0.75, 0.3, 0.0, -0.5):
                    cols, vals, n_hits, U = ix.search_sequence(seq.encode(), k, h, thr)
                    assert U == len(uk), (k, seq[:16], U, len(uk))
                    mk = max(math.ceil(len(uk) * thr), 0)
                    exp = np.nonzero(cnt >= mk)[0]
                    assert n_hits == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp]), (k, thr)
            # shorter than k: no window
            cols, vals, n_hits, U = ix.search_sequence(base[: k - 1].encode(), k, h, 1.0)
            assert U == 0 and n_hits == 0
        # a long sequence (beyond the in-kernel hashing limit): 300 kbp with a duplicated 50 kbp block
        k = 31
        oix = O.OracleIndex(k, m, h, N, rows=packed)
        long_seq = "".join(rng.choice(list("ACGT"), size=250_000))
        long_seq = long_seq + long_seq[40_000:90_000]
        uk = O.unique_kmers(long_seq, k)
        cnt = oix.counts(uk)
        mk = math.ceil(len(uk) * 0.6)
        cols, vals, n_hits, U = ix.search_sequence(long_seq.encode(), k, h, 0.6)
        exp = np.nonzero(cnt >= mk)[0]
        assert U == len(uk) and n_hits == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp])
    finally:
        ix.close()


def test_search_sequences_bulk_and_tickets(B):
    """bulk_search (one C-ABI call, pipelined searches with the front-end inside the gather kernel) and the
    submit / wait halves: every result identical to the oracle's search of the same sequence, whatever is in flight
    around it -- sequences with heavy repeats (many windows lose their table entry to another CTA's window),
    non-ACGT bytes, one window, no window, and a sequence too long for the streamed plan in the middle."""
    rng = np.random.default_rng(71)
    m, N, h, k = 20_011, 5000, 3, 31
    ix, packed = _random_index(B, rng, m, N, density=0.9)
    oix = O.OracleIndex(k, m, h, N, rows=packed)
    try:
        base = "".join(rng.choice(list("ACGT"), size=12_000))
        unit = base[:31]
        seqs = [base[:9000], base[100:400] * 20, unit * 28 + unit[:5], "ACGT" * 500, base[:k], base[: k - 1], "",
                base[2000:5000].lower() + "N" + base[:3000], base[:6000] + base[:6000], "A" * 5000,
                "".join(rng.choice(list("ACGT"), size=230_000)), base[500:9500], base[:4000] + "T" + base[:4000]]
        seqs = seqs + [base[i * 37 : i * 37 + 3000 + 13 * i] for i in range(30)]
        for thr in (1.0, 0.6, 0.0):
            res = ix.search_sequences([s.encode() for s in seqs], k, h, thr, cap=N)
            assert len(res) == len(seqs)
            for q, (seq, (cols, vals, n_hits, U)) in enumerate(zip(seqs, res)):
                uk = O.unique_kmers(seq, k)
                assert U == len(uk), (q, thr, U, len(uk))
                if not uk:
                    assert n_hits == 0
                    continue
                cnt = oix.counts(uk)
                exp = np.nonzero(cnt >= max(math.ceil(len(uk) * thr), 0))[0]
                assert n_hits == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp]), (q, thr)
        # a small capacity cuts the list, the count stays exact
        res = ix.search_sequences([s.encode() for s in seqs[:4]], k, h, 0.0, cap=16)
        assert all(n == N and len(c) == 16 for c, v, n, u in res)
        # tickets collected out of order, eight in flight
        tickets = [ix.search_sequence_submit(s.encode(), k, h, 0.7) for s in seqs[13:21]]
        with pytest.raises(B.BigsiB200Error):
            ix.search_sequence_submit(seqs[0].encode(), k, h, 0.7)  # a ninth
        for j in (3, 0, 7, 1, 2, 6, 5, 4):
            cols, vals, n_hits, U = ix.search_sequence_wait(tickets[j])
            uk = O.unique_kmers(seqs[13 + j], k)
            cnt = oix.counts(uk)
            exp = np.nonzero(cnt >= math.ceil(len(uk) * 0.7))[0]
            assert U == len(uk) and n_hits == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp]), j
        with pytest.raises(B.BigsiB200Error):
            ix.search_sequence_wait(tickets[0])  # collected already
    finally:
        ix.close()
