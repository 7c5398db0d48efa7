"""GPU parity tests of the rows next to the search path, through the C ABI:
Bloom construction + bit-transpose build (rank 3), score=True (rank 4), index files and the
reference's key/value schema (rank 2).  Bit-exact against the oracle and the golden vectors
generated from the unmodified reference."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden_util import bloom_from_b64, load
from tests.test_scoring_kv_golden import _oracle_search_with_score, kv_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import bigsi_b200

    assert bigsi_b200.device_count() >= 1, "no CUDA device visible to libbigsi_b200.so"
    return bigsi_b200


def _oracle_rows(blooms, m, n_bits=None):
    n_bits = m if n_bits is None else n_bits
    X = np.zeros((len(blooms), m), dtype=np.uint8)
    for i, b in enumerate(blooms):
        X[i, :n_bits] = np.unpackbits(b)[:n_bits]
    return np.packbits(X.T, axis=1) if len(blooms) else np.zeros((m, 0), dtype=np.uint8)


# ---------------------------------------------------------------------------
# K9: Bloom filters on the device
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("k,m,h,n", [(31, 25_000_000, 3, 5000), (31, 1000, 3, 100), (3, 25, 3, 1), (13, 99_991, 5, 777),
                                     (31, 257, 1, 64), (40, 4099, 2, 300)])
def test_bloom_kmers_matches_oracle(B, k, m, h, n):
    rng = np.random.default_rng(k * 1000 + n)
    arr = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, k))]
    kmers = [bytes(r).decode() for r in arr]
    want = O.OracleIndex.bloom(k, m, h, [O.canonical(x) for x in kmers])
    got = B.index.bloom_kmers(arr, k, h, m, canonical=True)
    assert np.array_equal(got, want)
    # canonical=False hashes the bytes as given (bare generate_hashes, bloom/bloomfilter.py:9-13)
    raw_bits = np.zeros(m, dtype=np.uint8)
    for km in kmers[:500]:
        raw_bits[list(O.generate_hashes(km, h, m))] = 1
    assert np.array_equal(B.index.bloom_kmers(arr[:500], k, h, m, canonical=False), np.packbits(raw_bits))
    # the BIGSI.bloom classmethod (graph/bigsi.py:150-155) on top of it
    cfg = {"k": k, "m": m, "h": h}
    assert np.array_equal(np.frombuffer(B.BIGSI.bloom(cfg, kmers).tobytes(), dtype=np.uint8), want)
    assert not np.any(B.index.bloom_kmers(arr[:0], k, h, m))


# ---------------------------------------------------------------------------
# K10: N x m -> m x N bit transpose
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("m,n", [(1000, 1), (1000, 33), (257, 31), (4099, 32), (999, 257), (70_001, 1000), (256, 64),
                                 (12_345, 8), (1, 5)])
def test_build_columns_matches_oracle_transpose(B, m, n):
    rng = np.random.default_rng(m + n)
    blooms = rng.integers(0, 256, size=(n, (m + 7) // 8), dtype=np.uint8)
    ix = B.DeviceIndex(m, 0, col_capacity=n)
    ix.build_columns(0, blooms, m)
    assert ix.num_cols == n
    assert np.array_equal(ix.download_rows(0, m), _oracle_rows(blooms, m))
    # the padding of the pitch stays zero (the search kernels rely on it)
    pitch = ix.info()["row_pitch_bytes"]
    ix.close()
    assert pitch % 128 == 0


def test_build_columns_appends_and_preserves(B):
    """Bulk insert: unaligned col0, several calls, filters shorter than m (rows >= n_bits get 0),
    overwrite of a middle range; every other column keeps its bits."""
    rng = np.random.default_rng(77)
    m = 3001
    nb = (m + 7) // 8
    parts = [rng.integers(0, 256, size=(c, nb), dtype=np.uint8) for c in (5, 30, 1, 64, 13)]
    ix = B.DeviceIndex(m, 0, col_capacity=200)
    col = 0
    for p in parts:
        ix.build_columns(col, p, m)
        col += p.shape[0]
    allb = np.concatenate(parts)
    assert ix.num_cols == col == 113
    assert np.array_equal(ix.download_rows(0, m), _oracle_rows(allb, m))
    # overwrite columns [17, 17+40) with short filters (2000 bits)
    repl = rng.integers(0, 256, size=(40, 250), dtype=np.uint8)
    ix.build_columns(17, repl, 2000)
    X = np.unpackbits(_oracle_rows(allb, m), axis=1)[:, :113]
    R = np.zeros((m, 40), dtype=np.uint8)
    R[:2000] = np.unpackbits(repl, axis=1)[:, :2000].T
    X[:, 17:57] = R
    assert np.array_equal(ix.download_rows(0, m), np.packbits(X, axis=1))
    # set_column (one-column insert) agrees with a one-filter build_columns
    one = rng.integers(0, 256, size=(1, nb), dtype=np.uint8)
    ix.set_column(113, one[0], m)
    ix2 = B.DeviceIndex(m, 0, col_capacity=200)
    ix2.build_columns(0, np.concatenate([np.packbits(X.T, axis=1)[:, :nb], one]), m)
    assert np.array_equal(ix.download_rows(0, m), ix2.download_rows(0, m))
    with pytest.raises(B.BigsiB200Error):
        ix.build_columns(150, parts[0], m)  # a gap: col0 beyond num_cols
    with pytest.raises(B.BigsiB200Error):
        ix.build_columns(114, rng.integers(0, 256, size=(2000, nb), dtype=np.uint8), m)  # beyond the capacity
    ix.close()
    ix2.close()


def test_build_columns_dev_from_device_filters(B):
    import torch

    rng = np.random.default_rng(5)
    m, n = 10_000, 70
    stride = (m + 255) // 256 * 32
    host = np.zeros((n, stride), dtype=np.uint8)
    host[:, : (m + 7) // 8] = rng.integers(0, 256, size=(n, (m + 7) // 8), dtype=np.uint8)
    d = torch.from_numpy(host).cuda()
    ix = B.DeviceIndex(m, 0, col_capacity=n)
    ix.build_columns_dev(0, n, d.data_ptr(), stride, m, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(ix.download_rows(0, m), _oracle_rows(host[:, : (m + 7) // 8], m))
    ix.close()


def test_build_then_search_at_scale(B):
    """A 2 000-sample index built by the transpose kernel from GPU-made Bloom filters answers like the oracle."""
    rng = np.random.default_rng(11)
    k, m, h, n = 31, 200_003, 3, 2000
    cfg = {"k": k, "m": m, "h": h, "storage-config": {"filename": "build-scale"}}
    base = "".join(rng.choice(list("ACGT"), size=3000))
    genomes = []
    for i in range(n):
        g = list(base[(i % 7) * 100 : (i % 7) * 100 + 1500])
        for p in rng.integers(0, len(g), size=i % 5):
            g[p] = "ACGT"[(("ACGT".index(g[p])) + 1) % 4]
        genomes.append("".join(g))
    names = ["g%d" % i for i in range(n)]
    blooms = [B.BIGSI.bloom(cfg, B.seq_to_kmers(g, k)) for g in genomes]
    bigsi = B.BIGSI.build(cfg, blooms, names)
    oix = O.OracleIndex.build(k, m, h, [np.frombuffer(b.tobytes(), dtype=np.uint8) for b in blooms], names)
    assert np.array_equal(bigsi.index.download_rows(0, m), oix.rows)
    for q, thr in ((base[:400], 1.0), (base[650:1200], 0.8), (genomes[3][:300], 0.95), (base[2000:2300], 0.0)):
        assert bigsi.search(q, thr) == oix.search(q, thr)
    bigsi.delete()


# ---------------------------------------------------------------------------
# K11 + Scorer: score=True
# ---------------------------------------------------------------------------
def test_search_with_score_golden(B):
    for ci, case in enumerate(load("scores.json")["searches"]):
        cfg = {"k": case["k"], "m": case["m"], "h": case["h"], "storage-config": {"filename": "score-%d" % ci}}
        blooms = []
        for seq, ref in zip(case["sample_seqs"], case["blooms_b64"]):
            b = B.BIGSI.bloom(cfg, B.seq_to_kmers(seq, case["k"]))
            assert np.array_equal(np.frombuffer(b.tobytes(), dtype=np.uint8), bloom_from_b64(ref))
            blooms.append(b)
        bigsi = B.BIGSI.build(cfg, blooms, case["samples"])
        for q in case["queries"]:
            if "raises" in q:
                with pytest.raises(BaseException) as ei:
                    bigsi.search(q["seq"], q["threshold"], score=True)
                assert type(ei.value).__name__ == q["raises"]
            else:
                got = bigsi.search(q["seq"], q["threshold"], score=True)
                assert got == q["result"], (q["seq"][:20], q["threshold"])
                if got:
                    assert list(got[0].keys()) == list(q["result"][0].keys())
        bigsi.delete()


def test_sequence_presence_vs_oracle(B):
    rng = np.random.default_rng(3)
    k, m, h, n = 31, 50_021, 3, 300
    ix = B.DeviceIndex(m, n)
    planted = [0, 7, 8, 150, n - 1]
    thr = [0xFFFFFFFF, 1 << 31, 3 << 30, 0xFFFFFFFF, 1 << 30]
    ix.fill_synthetic(9, 1, planted, thr)
    oix = O.OracleIndex(k, m, h, n, synth=O.SynthSpec(9, 1, planted, thr))
    seq = "".join(rng.choice(list("ACGT"), size=700))
    seq = seq + seq[100:300] + "N" + seq[:50]  # repeated windows and a non-ACGT byte
    cols = [0, 7, 8, 9, 150, 151, n - 1, 5]
    got = ix.sequence_presence(seq.encode(), k, h, cols)
    want = oix.presence_strings(seq, cols)
    assert got.shape == (len(cols), len(seq) - k + 1)
    assert [r.tobytes().decode() for r in got] == want
    assert ix.sequence_presence(seq[: k - 1].encode(), k, h, cols).shape == (len(cols), 0)
    with pytest.raises(B.BigsiB200Error):
        ix.sequence_presence(seq.encode(), k, h, [10 ** 6])
    ix.close()


# ---------------------------------------------------------------------------
# persistence: index files, the reference's key/value schema
# ---------------------------------------------------------------------------
def test_save_load_round_trip(B, tmp_path):
    rng = np.random.default_rng(21)
    k, m, h, n = 21, 40_009, 3, 333
    cfg = {"k": k, "m": m, "h": h, "storage-config": {"filename": "persist-a"}}
    genomes = ["".join(rng.choice(list("ACGT"), size=400)) for _ in range(n)]
    names = ["p%d" % i for i in range(n)]
    bigsi = B.BIGSI.build(cfg, [B.BIGSI.bloom(cfg, B.seq_to_kmers(g, k)) for g in genomes], names)
    bigsi.delete_sample("p5")
    rows = bigsi.index.download_rows(0, m)
    queries = [(genomes[0][:120], 1.0), (genomes[5][:90], 1.0), (genomes[7][50:200], 0.5), (genomes[1][:60], 0.0)]
    want = [bigsi.search(q, t) for q, t in queries]
    path = str(tmp_path / "index.bigsib2")
    bigsi.save(path)
    hd, meta = B.index.file_info(path)
    assert (hd["num_rows"], hd["num_cols"], hd["row_bytes"], hd["version"]) == (m, n, (n + 7) // 8, 1)
    assert hd["rows_offset"] % 4096 == 0 and os.path.getsize(path) == hd["rows_offset"] + m * hd["row_bytes"]
    with open(path, "rb") as f:  # the row region IS the concatenation of the reference's row values
        f.seek(hd["rows_offset"])
        assert f.read() == rows.tobytes()
    bigsi.delete()
    cfg2 = {"storage-config": {"filename": "persist-b"}}
    re = B.BIGSI.load(cfg2, path)
    assert (re.kmer_size, re.bloomfilter_size, re.num_hashes, re.num_samples) == (k, m, h, n)
    assert np.array_equal(re.index.download_rows(0, m), rows)
    assert [re.search(q, t) for q, t in queries] == want
    assert re.colour_to_sample(5) == B.DELETION_SPECIAL_SAMPLE_NAME and re.sample_to_colour("p6") == 6
    # a column shard loads its byte range of the full-width file
    shard = B.DeviceIndex(m, 100, col_offset=96)
    shard.load_rows(path, hd["rows_offset"], hd["row_bytes"], src_byte_offset=12, row0=0, n_rows=m)
    want_shard = rows[:, 12:25].copy()
    want_shard[:, -1] &= 0xF0  # 100 columns: the last 4 bits of byte 12 are padding
    assert np.array_equal(shard.download_rows(0, m), want_shard)
    shard.close()
    re.delete()
    with pytest.raises(B.BigsiB200Error):
        B.index.file_info(str(tmp_path / "missing"))
    bad = tmp_path / "bad"
    bad.write_bytes(b"not an index" * 10)
    with pytest.raises(B.BigsiB200Error):
        B.index.file_info(str(bad))


def test_reference_kv_schema_import_export(B):
    """from_kv on the reference's own store content answers the golden queries; to_kv writes the very
    same key/value pairs back (every key, every byte) -- a reference backend could be filled from it."""
    for ci, case in enumerate(load("kv_store.json")):
        kv = kv_of(case)
        cfg = {"k": case["k"], "storage-config": {"filename": "kv-%d" % ci}}
        bigsi = B.BIGSI.from_kv(cfg, kv)
        assert (bigsi.bloomfilter_size, bigsi.num_hashes, bigsi.num_samples) == (case["m"], case["h"], case["num_samples"])
        for q in case["queries"]:
            assert bigsi.search(q["seq"], q["threshold"]) == q["result"]
        assert bigsi.to_kv() == kv
        if case["deleted"]:
            assert bigsi.sample_to_colour(case["deleted"]) is None
        bigsi.delete()
    # and the other way round: an index built here exports what the reference would have stored
    case = load("kv_store.json")[2]  # built without insert/delete
    cfg = {"k": case["k"], "m": case["m"], "h": case["h"], "storage-config": {"filename": "kv-export"}}
    bigsi = B.BIGSI.build(cfg, [B.BIGSI.bloom(cfg, B.seq_to_kmers(s, case["k"])) for s in case["sample_seqs"]], case["samples"])
    assert bigsi.to_kv() == kv_of(case)
    bigsi.delete()


# ---------------------------------------------------------------------------
# BIGSI(config) column-sharded over several GPUs (storage-config.devices, bigsi_b200/sharded_index.py)
# ---------------------------------------------------------------------------
def test_sharded_bigsi_matches_single_gpu(B, tmp_path):
    """build / insert / search (exact, inexact, score=True) / lookup / delete_sample / to_kv / save / load through a
    column-sharded index give what the single-GPU index gives, bit for bit: three shards (distinct GPUs where the box
    has them), appends that land in the last shard and grow it, a file written sharded and loaded on one GPU and on two
    shards."""
    import torch

    ndev = max(torch.cuda.device_count(), 1)
    devs = [0, 1 % ndev, 2 % ndev]
    rng = np.random.default_rng(77)
    k, m, h, n = 15, 30_011, 3, 45
    genomes = ["".join(rng.choice(list("ACGT"), size=500)) for _ in range(n + 4)]
    for i in range(2, n, 4):
        genomes[i] = genomes[i - 1][:300] + genomes[i][300:]
    names = ["g%d" % i for i in range(n + 4)]
    one_cfg = {"k": k, "m": m, "h": h, "storage-config": {"filename": "sh-one", "device": 0}}
    sh_cfg = {"k": k, "m": m, "h": h, "storage-config": {"filename": "sh-three", "devices": devs}}
    blooms = [B.BIGSI.bloom(one_cfg, B.seq_to_kmers(g, k)) for g in genomes]
    one = B.BIGSI.build(one_cfg, blooms[:n], names[:n])
    sh = B.BIGSI.build(sh_cfg, blooms[:n], names[:n])
    info = sh.index.info()
    assert [i["num_cols"] for i in info["shards"]] == [16, 16, 13] and [i["col_offset"] for i in info["shards"]] == [0, 16, 32]
    queries = [(genomes[0][:150], 1.0), (genomes[1][200:420], 0.8), (genomes[17][:80] + "N" + genomes[40][:80], 0.3),
               (genomes[44][:k], 1.0), (genomes[3][:200], 0.0), ("ACGT" * 30, 1.0)]

    def same(a, b):
        assert np.array_equal(a.index.download_rows(0, m), b.index.download_rows(0, m))
        for q, t in queries:
            assert a.search(q, t) == b.search(q, t), (q[:20], t)
            if t < 1.0:
                assert a.search(q, t, score=True) == b.search(q, t, score=True)
        kms = list(B.seq_to_kmers(genomes[2][:60], k))
        la, lb = a.lookup(kms), b.lookup(kms)
        assert {x: v.to01() for x, v in la.items()} == {x: v.to01() for x, v in lb.items()}
        assert a.to_kv() == b.to_kv()

    same(one, sh)
    for j in range(n, n + 4):  # appends: the last shard takes them
        one.insert(blooms[j], names[j])
        sh.insert(blooms[j], names[j])
    assert sh.index.num_cols == n + 4 and sh.num_samples == n + 4
    one.delete_sample("g7")
    sh.delete_sample("g7")
    queries.append((genomes[n + 2][:140], 1.0))
    same(one, sh)
    path = str(tmp_path / "sharded.bigsib2")
    sh.save(path)
    single_path = str(tmp_path / "single.bigsib2")
    one.save(single_path)
    with open(path, "rb") as f1, open(single_path, "rb") as f2:  # one file format, whatever wrote it
        assert f1.read() == f2.read()
    re1 = B.BIGSI.load({"storage-config": {"filename": "sh-re1", "device": 0}}, path)
    re2 = B.BIGSI.load({"storage-config": {"filename": "sh-re2", "devices": devs[:2]}}, path)
    # a BIGSI object fixes its Scorer's database size at construction (graph/bigsi.py:140), so `one` (45 samples when it
    # was built, 49 now) is compared through a fresh handle on the same resident store, like the freshly loaded ones
    one = B.BIGSI(one_cfg)
    same(one, re1)
    same(one, re2)
    for b in (one, sh, re1, re2):
        b.delete()
    with pytest.raises(BaseException):
        B.BIGSI(sh_cfg)


def test_sharded_append_crosses_shards(B):
    """Fewer samples than shards: the later shards start empty; appends fill the first shard's pitch (1 024 columns),
    then open the next shard exactly there."""
    rng = np.random.default_rng(79)
    k, m, h = 9, 997, 2
    cfg = {"k": k, "m": m, "h": h, "storage-config": {"filename": "sh-append", "devices": [0, 0]}}
    one_cfg = {"k": k, "m": m, "h": h, "storage-config": {"filename": "sh-append-one", "device": 0}}
    seqs = ["".join(rng.choice(list("ACGT"), size=40)) for _ in range(6)]
    blooms = [B.BIGSI.bloom(cfg, B.seq_to_kmers(s, k)) for s in seqs]
    sh = B.BIGSI.build(cfg, blooms[:3], ["a", "b", "c"])
    one = B.BIGSI.build(one_cfg, blooms[:3], ["a", "b", "c"])
    assert [i["num_cols"] for i in sh.index.info()["shards"]] == [3, 0]
    filler = B.BIGSI.bloom(cfg, B.seq_to_kmers(seqs[3], k))
    for j in range(3, 1027):
        b = blooms[4] if j == 1025 else filler
        sh.insert(b, "x%d" % j)
        one.insert(b, "x%d" % j)
    info = sh.index.info()
    assert [i["num_cols"] for i in info["shards"]] == [1024, 3] and info["shards"][1]["col_offset"] == 1024
    for q, t in ((seqs[4][:30], 1.0), (seqs[0], 1.0), (seqs[3][:20], 0.5)):
        assert sh.search(q, t) == one.search(q, t)
    assert np.array_equal(sh.index.download_rows(0, m), one.index.download_rows(0, m))
    sh.delete()
    one.delete()
