"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

Binds oracle/bigsi_oracle.c (built by oracle/Makefile into
oracle/libbigsi_oracle.so) with ctypes and restates the host-side logic of the
reference's search path (unique raw k-mers, min_kmers, result dicts, ordering).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; bigsi_b200/ never does.

Reference anchors (all under /root/reference/bigsi/):
  utils/fncts.py:63-65      seq_to_kmers
  graph/index.py:42-80      lookup (set of raw k-mers, hashes, AND per k-mer)
  graph/bigsi.py:174-230    search / exact_filter / inexact_filter
  graph/bigsi.py:91-126     BigsiQueryResult dict layout and rounding
  graph/metadata.py:1       tombstone sample name
Parity pin: tests/golden/*.json (generated from the unmodified reference by
tests/golden/make_golden.py) -- see tests/test_oracle_golden.py.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbigsi_oracle.so")
DELETION_SPECIAL_SAMPLE_NAME = "D3L3T3D"  # graph/metadata.py:1

_lib = None


def build(force=False):
    """Compile the C restatement (gcc -O3 -fopenmp)."""
    src = os.path.join(_HERE, "bigsi_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libbigsi_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.oracle_murmur3_x86_32.restype = ctypes.c_uint32
        L.oracle_murmur3_x86_32.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32]
        L.oracle_hash_row.restype = ctypes.c_int64
        L.oracle_hash_row.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_int64]
        L.oracle_canonical.restype = None
        L.oracle_canonical.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p]
        L.oracle_hash_kmers.restype = None
        L.oracle_hash_kmers.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int64, ctypes.c_void_p]
        for name in ("oracle_and_per_kmer", "oracle_exact", "oracle_counts"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                          ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        L.oracle_synth_rows.restype = None
        L.oracle_synth_rows.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------
def murmur3_32_signed(key, seed=0):
    """mmh3.hash(key, seed) (signed)."""
    if isinstance(key, str):
        key = key.encode("utf-8")
    v = lib().oracle_murmur3_x86_32(key, len(key), seed & 0xFFFFFFFF)
    return v - (1 << 32) if v >= (1 << 31) else v


def canonical(kmer):
    """utils/fncts.py:51-54."""
    b = kmer.encode("utf-8") if isinstance(kmer, str) else bytes(kmer)
    out = ctypes.create_string_buffer(len(b))
    lib().oracle_canonical(b, len(b), out)
    r = out.raw[: len(b)]
    return r.decode("utf-8") if isinstance(kmer, str) else r


def generate_hashes(element, h, m):
    """bloom/bloomfilter.py:9-13 (NO canonicalisation here, as in the reference)."""
    b = element.encode("utf-8") if isinstance(element, str) else bytes(element)
    return {int(lib().oracle_hash_row(b, len(b), s, m)) for s in range(h)}


def kmers_to_array(kmers, k):
    """list of equal-length k-mer strings -> uint8 [U, k]."""
    if len(kmers) == 0:
        return np.zeros((0, k), dtype=np.uint8)
    joined = "".join(kmers).encode("utf-8")
    assert len(joined) == len(kmers) * k, "k-mers must be ASCII and of length k"
    return np.frombuffer(joined, dtype=np.uint8).reshape(len(kmers), k).copy()


def hash_kmers(kmers, k, h, m):
    """Canonical + h signed-murmur3 floor-mod row ids; int32 [U, h] (graph/index.py:62-70)."""
    arr = kmers if isinstance(kmers, np.ndarray) else kmers_to_array(kmers, k)
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    out = np.empty((arr.shape[0], h), dtype=np.int32)
    lib().oracle_hash_kmers(_ptr(arr), arr.shape[0], k, h, m, _ptr(out))
    return out


def seq_to_kmers(seq, k):
    """utils/fncts.py:63-65."""
    return [seq[i : i + k] for i in range(len(seq) - k + 1)]


def unique_kmers(seq_or_kmers, k):
    """set(kmers) of graph/index.py:45 with a deterministic (first-occurrence) order."""
    kmers = seq_to_kmers(seq_or_kmers, k) if isinstance(seq_or_kmers, str) else list(seq_or_kmers)
    return list(dict.fromkeys(kmers))


# ---------------------------------------------------------------------------
# synthetic matrix rows (same pure function as the device fill kernel)
# ---------------------------------------------------------------------------
class SynthSpec:
    def __init__(self, seed=0, and_draws=1, planted_cols=(), planted_thr=()):
        self.seed = int(seed)
        self.and_draws = int(and_draws)
        self.planted_cols = np.asarray(planted_cols, dtype=np.uint64)
        self.planted_thr = np.asarray(planted_thr, dtype=np.uint32)
        assert self.planted_cols.shape == self.planted_thr.shape

    def rows(self, row_ids, col_offset, num_cols, stride=None):
        row_ids = np.ascontiguousarray(row_ids, dtype=np.int64)
        row_bytes = (num_cols + 7) // 8
        stride = stride or row_bytes
        out = np.empty((len(row_ids), stride), dtype=np.uint8)
        lib().oracle_synth_rows(self.seed, self.and_draws, _ptr(row_ids), len(row_ids), col_offset,
                                num_cols, row_bytes, stride, _ptr(self.planted_cols),
                                _ptr(self.planted_thr), len(self.planted_cols), _ptr(out))
        return out


# ---------------------------------------------------------------------------
# index restatement
# ---------------------------------------------------------------------------
class OracleIndex:
    """CPU restatement of KmerSignatureIndex + BIGSI.search over packed MSB-first rows.

    Either `rows` (dense uint8 [m, ceil(N/8)]) or `synth` (SynthSpec; rows regenerated
    on demand, only the touched ones) backs the matrix.
    """

    def __init__(self, k, m, h, num_cols, rows=None, synth=None, samples=None, col_offset=0):
        self.k, self.m, self.h, self.num_cols = k, m, h, num_cols
        self.row_bytes = (num_cols + 7) // 8
        self.rows = rows
        self.synth = synth
        self.col_offset = col_offset
        self.samples = list(samples) if samples is not None else [str(i) for i in range(num_cols)]
        if rows is not None:
            assert rows.dtype == np.uint8 and rows.shape == (m, self.row_bytes)

    # -- build path (graph/bigsi.py:150-172, matrix/transpose.py:33-43) ------
    @staticmethod
    def bloom(k, m, h, kmers):
        """BIGSI.bloom: canonical k-mers -> Bloom filter, packed MSB-first uint8[ceil(m/8)]."""
        bits = np.zeros(m, dtype=np.uint8)
        kmers = list(kmers)
        if kmers:
            klen = len(kmers[0])
            r = hash_kmers(kmers, klen, h, m)
            bits[r.reshape(-1)] = 1
        return np.packbits(bits)

    @classmethod
    def build(cls, k, m, h, blooms, samples):
        if len(blooms) != len(samples):
            raise ValueError("There must be the same number of bloomfilters and sample names")
        X = np.stack([np.unpackbits(b)[:m] for b in blooms], axis=0)  # [N, m]
        rows = np.packbits(X.T, axis=1)  # [m, ceil(N/8)]
        return cls(k, m, h, len(blooms), rows=np.ascontiguousarray(rows), samples=samples)

    # -- row access ----------------------------------------------------------
    def _store(self, row_ids_2d):
        """Returns (store uint8 [R, row_bytes], slot int64 [U, h])."""
        if self.rows is not None:
            return self.rows, np.ascontiguousarray(row_ids_2d, dtype=np.int64)
        uniq, inv = np.unique(row_ids_2d.reshape(-1), return_inverse=True)
        store = self.synth.rows(uniq, self.col_offset, self.num_cols)
        return store, np.ascontiguousarray(inv.reshape(row_ids_2d.shape), dtype=np.int64)

    # -- graph/index.py:42-49 -------------------------------------------------
    def lookup_packed(self, kmers):
        """unique raw k-mers (list) -> uint8 [U, row_bytes] per-k-mer AND vectors."""
        r = hash_kmers(kmers, self.k, self.h, self.m)
        store, slot = self._store(r)
        out = np.empty((len(kmers), self.row_bytes), dtype=np.uint8)
        if len(kmers):
            lib().oracle_and_per_kmer(_ptr(store), store.strides[0], self.row_bytes, _ptr(slot),
                                      len(kmers), self.h, _ptr(out))
        return out

    def lookup(self, kmers):
        """{raw kmer: '0101..' string of num_cols bits} (remove_trailing_zeros=True view)."""
        if isinstance(kmers, str):
            kmers = [kmers]
        uk = unique_kmers(kmers, self.k)
        packed = self.lookup_packed(uk)
        return {km: "".join(map(str, np.unpackbits(packed[i])[: self.num_cols])) for i, km in enumerate(uk)}

    def presence(self, kmers):
        """exact_filter's AND over all k-mers: packed uint8 [row_bytes]."""
        r = hash_kmers(kmers, self.k, self.h, self.m)
        store, slot = self._store(r)
        out = np.empty(self.row_bytes, dtype=np.uint8)
        lib().oracle_exact(_ptr(store), store.strides[0], self.row_bytes, _ptr(slot), len(kmers), self.h, _ptr(out))
        return out

    def counts(self, kmers):
        """unpack_and_sum truncated to num_cols: int32 [N]."""
        r = hash_kmers(kmers, self.k, self.h, self.m)
        store, slot = self._store(r)
        out = np.empty(self.row_bytes * 8, dtype=np.int32)
        lib().oracle_counts(_ptr(store), store.strides[0], self.row_bytes, _ptr(slot), len(kmers), self.h, _ptr(out))
        return out[: self.num_cols]

    # -- graph/bigsi.py:174-230 ----------------------------------------------
    def search(self, seq, threshold=1.0):
        assert threshold <= 1
        uk = unique_kmers(seq, self.k)
        U = len(uk)
        if U == 0:
            # reduce() of an empty sequence at utils/fncts.py:24-25
            raise TypeError("reduce() of empty iterable with no initial value")
        min_kmers = math.ceil(U * threshold)
        if threshold == 1.0:
            pres = np.unpackbits(self.presence(uk))[: self.num_cols]
            hits = [(int(c), U) for c in np.nonzero(pres)[0]]
        else:
            cnt = self.counts(uk)
            cols = np.nonzero(cnt >= min_kmers)[0]
            order = np.argsort(-cnt[cols], kind="stable")
            hits = [(int(cols[i]), int(cnt[cols[i]])) for i in order]
        return format_results(hits, U, self.samples)


    # -- graph/bigsi.py:232-239 (score=True): per-window presence of a column -----
    def presence_strings(self, seq, cols):
        """For every window of seq (sequence order, duplicates included) and every column in cols:
        '1' if the window's k-mer is present in the column -- what the reference extracts with
        unpack_and_cat (graph/bigsi.py:47-56) as X[:, colour]."""
        kmers = list(seq_to_kmers(seq, self.k))
        if not kmers:
            return ["" for _ in cols]
        uk = unique_kmers(kmers, self.k)
        bits = np.unpackbits(self.lookup_packed(uk), axis=1)  # [U, 8*row_bytes]
        pos = {km: i for i, km in enumerate(uk)}
        order = np.array([pos[km] for km in kmers], dtype=np.int64)
        return ["".join("1" if b else "0" for b in bits[order, c]) for c in cols]

    # -- storage schema (storage/base.py:29-52,77-94) ------
    def to_kv(self, metadata):
        """The reference's v0.3 key/value store content for this index.  metadata: dict with
        "colour_count", {sample name: colour} under "samples" and {colour: name} under "colours"."""
        assert self.rows is not None
        kv = {}

        def put_int(key, v):
            kv[(key + ":int").encode()] = str(int(v)).encode()

        put_int("ksi:bloomfilter_size", self.m)
        put_int("ksi:num_hashes", self.h)
        put_int("number_of_rows", self.m)
        put_int("number_of_cols", self.num_cols)
        put_int("metadata:colour_count", metadata["colour_count"])
        for name, colour in metadata["samples"].items():
            put_int("metadata:%s" % name, colour)
        for colour, name in metadata["colours"].items():
            kv[("metadata:%d:string" % colour).encode()] = name.encode()
        for r in range(self.m):
            kv[b"%d:bitarray" % r] = self.rows[r].tobytes()
        return kv


# ---------------------------------------------------------------------------
# row-id level entry points (used to check the device kernels without hashing)
# ---------------------------------------------------------------------------
def counts_from_rows(store, row_ids, num_cols):
    """store uint8 [m, row_bytes]; row_ids int [U, h] -> int32 [num_cols] (graph/bigsi.py:35-44)."""
    store = np.ascontiguousarray(store, dtype=np.uint8)
    slot = np.ascontiguousarray(row_ids, dtype=np.int64)
    rb = store.shape[1]
    out = np.zeros(rb * 8, dtype=np.int32)
    if slot.shape[0]:
        lib().oracle_counts(_ptr(store), store.strides[0], rb, _ptr(slot), slot.shape[0], slot.shape[1], _ptr(out))
    return out[:num_cols]


def presence_from_rows(store, row_ids):
    """AND over all k-mers (graph/bigsi.py:192-195): packed uint8 [row_bytes]."""
    store = np.ascontiguousarray(store, dtype=np.uint8)
    slot = np.ascontiguousarray(row_ids, dtype=np.int64)
    rb = store.shape[1]
    out = np.empty(rb, dtype=np.uint8)
    lib().oracle_exact(_ptr(store), store.strides[0], rb, _ptr(slot), slot.shape[0], slot.shape[1], _ptr(out))
    return out


def and_per_kmer_from_rows(store, row_ids):
    """graph/index.py:75-80: uint8 [U, row_bytes]."""
    store = np.ascontiguousarray(store, dtype=np.uint8)
    slot = np.ascontiguousarray(row_ids, dtype=np.int64)
    rb = store.shape[1]
    out = np.empty((slot.shape[0], rb), dtype=np.uint8)
    if slot.shape[0]:
        lib().oracle_and_per_kmer(_ptr(store), store.strides[0], rb, _ptr(slot), slot.shape[0], slot.shape[1], _ptr(out))
    return out


def format_results(hits, num_kmers, samples):
    """graph/bigsi.py:91-126 + 186-190: dict layout, rounding, tombstone filter."""
    out = []
    for colour, found in hits:
        name = samples[colour]
        if name == DELETION_SPECIAL_SAMPLE_NAME:
            continue
        out.append({
            "percent_kmers_found": round(100 * float(found) / num_kmers, 2),
            "num_kmers": num_kmers,
            "num_kmers_found": found,
            "sample_name": name,
        })
    return out
