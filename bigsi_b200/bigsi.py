"""BIGSI: drop-in for the reference's public class on the search path.

Mirrors bigsi/graph/bigsi.py:129-275 (constructor from a config dict, classmethods bloom/build,
insert, search, lookup, delete, result dict layout) while the data plane is the HBM-resident
DeviceIndex + the fused CUDA kernel:

  reference                                              here
  ------------------------------------------------------ -------------------------------------------
  get_storage(config) (storage/__init__.py:18-19)        process-resident store keyed by config
  KmerSignatureIndex.lookup (graph/index.py:42-80)       bigsi_b200_lookup_kmers (hash + gather-AND)
  exact_filter (graph/bigsi.py:192-205)                  bigsi_b200_search_kmers  MODE_AND
  inexact_filter (graph/bigsi.py:211-230)                bigsi_b200_search_kmers_hits (counts + threshold)

The reference constructs BIGSI(config) per request/worker (bigsi/__main__.py:204,76); the GPU
index is therefore cached per process, keyed by the storage name, and never re-uploaded per call.
"""
import json
import logging
import math
import threading

import numpy as np

from . import bits as _bits
from ._lib import MODE_AND, MODE_COUNTS
from .bloom import BloomFilter
from .index import DeviceIndex, bloom_kmers, file_info, hash_kmers, kmers_to_array
from .sharded_index import make_index
from .metadata import DELETION_SPECIAL_SAMPLE_NAME, SampleMetadata
from .scoring import Scorer
from .utils import convert_query_kmers, seq_to_kmers, unique_kmers

logger = logging.getLogger(__name__)

DEFAULT_CONFIG = {  # bigsi/constants.py:13-19 with the HBM engine in place of a KV backend
    "h": 3,
    "k": 31,
    "m": 25 * 10 ** 6,
    "storage-engine": "b200",
    # "device": one GPU; "devices": [0, 1, ...] column-shards the matrix over several GPUs of the box (sharded_index.py)
    "storage-config": {"filename": "bigsi-b200-default", "device": 0},
}
DEFAULT_NPROC = 4
MIN_UNIQUE_KMERS_IN_QUERY = 0

_STORES = {}  # (storage name, devices) -> _Store (the process-resident replacement of the KV store)
_STORES_LOCK = threading.RLock()


class _Store:
    """What the reference keeps in its KV store: the bit matrix, ksi:* and metadata:* keys."""

    def __init__(self, index, m, h, k):
        self.index = index
        self.bloomfilter_size = m
        self.num_hashes = h
        self.kmer_size = k
        self.meta = {}

    def close(self):
        self.index.close()


def _store_name(config):
    sc = config.get("storage-config", {}) or {}
    return str(sc.get("filename", sc.get("name", "bigsi-b200-default")))


def _devices(config):
    """storage-config.devices (a list: column shards over several GPUs) or storage-config.device (one GPU)."""
    sc = config.get("storage-config", {}) or {}
    devs = sc.get("devices")
    if devs is None:
        return [int(sc.get("device", 0))]
    if isinstance(devs, int):
        devs = [devs]
    devs = [int(d) for d in devs]  # (an ordinal may repeat: several shards on one GPU -- pointless in production, handy for tests)
    if not devs:
        raise ValueError("storage-config.devices must be a non-empty list of device ordinals")
    return devs


def _device(config):
    return _devices(config)[0]


def _store_key(config):
    return (_store_name(config), tuple(_devices(config)))


def _register(config, store):
    """Put `store` under config's key; a store it replaces is closed (objects that still hold it get 'index has been
    destroyed' from their next call, as with a reference store deleted under its users)."""
    with _STORES_LOCK:
        old = _STORES.pop(_store_key(config), None)
        _STORES[_store_key(config)] = store
    if old is not None and old is not store:
        old.close()


def merge_packed_rows(a, n1, b, n2):
    """Bit-concatenate packed MSB-first rows: uint8 [n, >= ceil(n1/8)] and [n, >= ceil(n2/8)] ->
    uint8 [n, ceil((n1+n2)/8)]; columns 0..n1-1 from a, n1..n1+n2-1 from b (graph/index.py:54-60)."""
    n = a.shape[0]
    A = np.unpackbits(np.ascontiguousarray(a, dtype=np.uint8), axis=1)[:, :n1] if n1 else np.zeros((n, 0), dtype=np.uint8)
    Bb = np.unpackbits(np.ascontiguousarray(b, dtype=np.uint8), axis=1)[:, :n2] if n2 else np.zeros((n, 0), dtype=np.uint8)
    if n1 + n2 == 0:
        return np.zeros((n, 0), dtype=np.uint8)
    return np.packbits(np.concatenate([A, Bb], axis=1), axis=1)


def validate_build_params(bloomfilters, samples):
    if not len(bloomfilters) == len(samples):
        raise ValueError("There must be the same number of bloomfilters and sample names")


class BigsiQueryResult:
    """bigsi/graph/bigsi.py:91-126."""

    PERCENT_KMERS_FOUND_KEY = "percent_kmers_found"
    NUM_KMERS_KEY = "num_kmers"
    NUM_KMERS_FOUND_KEY = "num_kmers_found"
    SAMPLE_KEY = "sample_name"

    def __init__(self, colour, sample_name, num_kmers_found, num_kmers):
        self.colour = colour
        self.sample_name = sample_name
        self.num_kmers_found = num_kmers_found
        self.num_kmers = num_kmers
        self.percent_kmers_found = round(100 * float(num_kmers_found) / num_kmers, 2)
        self.score = None

    def todict(self):
        outd = {
            self.PERCENT_KMERS_FOUND_KEY: self.percent_kmers_found,
            self.NUM_KMERS_KEY: self.num_kmers,
            self.NUM_KMERS_FOUND_KEY: self.num_kmers_found,
            self.SAMPLE_KEY: self.sample_name,
        }
        if self.score:
            outd.update(self.score)
        return outd

    def __eq__(self, ob):
        return self.todict() == ob.todict()

    def add_score(self, score):
        self.score = score


class BIGSI(SampleMetadata):
    def __init__(self, config=None):
        if config is None:
            config = DEFAULT_CONFIG
        self.config = config
        key = _store_key(config)
        with _STORES_LOCK:
            if key not in _STORES:
                # the reference raises KeyError from storage.get_integer on an empty store
                # (graph/index.py:24, matrix/bitmatrix.py:16-17)
                raise KeyError("no index has been built for storage '%s' on devices %s" % key)
            self._store = _STORES[key]
        SampleMetadata.__init__(self, self._store.meta)
        self.min_unique_kmers_in_query = MIN_UNIQUE_KMERS_IN_QUERY
        self.scorer = Scorer(self.num_samples)  # graph/bigsi.py:140 (DB size fixed at construction)

    # -- properties ------------------------------------------------------------
    @property
    def index(self):
        return self._store.index

    @property
    def kmer_size(self):
        return self.config["k"]

    @property
    def nproc(self):
        return self.config.get("nproc", DEFAULT_NPROC)

    @property
    def bloomfilter_size(self):
        return self._store.bloomfilter_size

    @property
    def num_hashes(self):
        return self._store.num_hashes

    # -- build path ------------------------------------------------------------
    @classmethod
    def bloom(cls, config, kmers):
        """graph/bigsi.py:150-155: canonical k-mers -> Bloom filter bitarray (hashing on the GPU)."""
        kmers = convert_query_kmers(kmers)
        bloomfilter = BloomFilter(m=config["m"], h=config["h"], device=_device(config))
        bloomfilter.update(kmers)
        return bloomfilter.bitarray

    @classmethod
    def build(cls, config, bloomfilters, samples):
        """graph/bigsi.py:157-172: N Bloom filters become the N columns of the m x N matrix."""
        validate_build_params(bloomfilters, samples)
        m, h, k = config["m"], config["h"], config["k"]
        with _STORES_LOCK:
            old = _STORES.pop(_store_key(config), None)
        if old is not None:
            old.close()
        n = len(bloomfilters)
        sc = config.get("storage-config", {}) or {}
        capacity = int(sc.get("col_capacity", 0)) or max(n, 1)
        index = make_index(m, n, col_capacity=capacity, devices=_devices(config))
        store = _Store(index, m, h, k)
        try:
            SampleMetadata(store.meta).add_samples(samples)
            nbytes = (m + 7) // 8
            # matrix/transpose.py:33-50 + matrix/bitmatrix.py:19-25: the N filters of m bits become the N
            # columns of the m x N matrix -- one bit-transpose kernel per staged group of filters
            group = max(32, ((256 << 20) // max(nbytes, 1)) // 32 * 32)
            for c0 in range(0, n, group):
                packed = np.zeros((min(group, n - c0), nbytes), dtype=np.uint8)
                for i, bf in enumerate(bloomfilters[c0 : c0 + group]):
                    if isinstance(bf, BloomFilter):
                        bf = bf.bitarray
                    p = _bits.to_packed(bf, m)
                    if p.size * 8 < m:
                        raise ValueError("bloom filter shorter than m=%d bits" % m)
                    packed[i] = p[:nbytes]
                index.build_columns(c0, packed, m)
        except Exception:
            store.close()
            raise
        _register(config, store)
        return cls(config)

    def insert(self, bloomfilter, sample):
        """graph/bigsi.py:244-247 -> matrix/bitmatrix.py:67-75 (column = new count - 1)."""
        logger.warning("Build and merge is preferable to insert in most cases")
        colour = self.add_sample(sample)
        self.insert_bloom(bloomfilter, colour - 1)

    def insert_bloom(self, bloomfilter, column_index):
        if isinstance(bloomfilter, BloomFilter):
            bloomfilter = bloomfilter.bitarray
        m = self.bloomfilter_size
        p = _bits.to_packed(bloomfilter, m)
        nbits = _bits.nbits_of(bloomfilter)
        nbits = m if nbits is None else min(nbits, m)
        info = self.index.info()
        if column_index >= info["col_capacity"]:
            # re-pitch when an insert exceeds the column capacity (a sharded index grows its last shard itself)
            self._store.index = self.index.grown(max(column_index + 1, 2 * info["col_capacity"]))
        self.index.set_column(column_index, p, nbits)

    def delete(self):
        """graph/bigsi.py:249-250: storage.delete_all()."""
        with _STORES_LOCK:
            st = _STORES.pop(_store_key(self.config), None)
        if st is not None:
            st.close()

    def merge(self, bigsi):
        """graph/bigsi.py:252-260: append the columns and the samples of another index with the same
        (m, h, k).  Rows are cut to num_cols bits and extended (graph/index.py:54-60); samples are added in
        colour order, a name that cannot be added (duplicate, or the tombstone of a deleted sample) gets the
        suffix "_duplicate_in_merge" (graph/metadata.py:74-80).  An offline maintenance path: the rows make a
        host round trip in chunks (download, bit-concatenate, upload into a re-pitched matrix)."""
        assert self.bloomfilter_size == bigsi.bloomfilter_size
        assert self.num_hashes == bigsi.num_hashes
        assert self.kmer_size == bigsi.kmer_size
        old, other = self.index, bigsi.index
        info = old.info()
        n1, n2 = info["num_cols"], other.num_cols
        new = make_index(info["num_rows"], n1 + n2, col_capacity=max(info["col_capacity"], n1 + n2),
                         col_offset=info["col_offset"], devices=_devices(self.config))
        try:
            step = max(1, (1 << 24) // max((n1 + n2 + 7) // 8, 1))
            for r0 in range(0, info["num_rows"], step):
                n = min(step, info["num_rows"] - r0)
                new.upload_rows(r0, merge_packed_rows(old.download_rows(r0, n), n1, other.download_rows(r0, n), n2))
        except Exception:
            new.close()
            raise
        self._store.index = new
        old.close()
        for c in range(bigsi.num_samples):
            sample = bigsi.colour_to_sample(c)
            try:
                self.add_sample(sample)
            except ValueError:
                self.add_sample(sample + "_duplicate_in_merge")

    # -- query path ------------------------------------------------------------
    def seq_to_kmers(self, seq):
        return seq_to_kmers(seq, self.kmer_size)

    def lookup(self, kmers, remove_trailing_zeros=True):
        """graph/index.py:42-49: {raw k-mer: bitarray of the samples that contain it}."""
        if isinstance(kmers, str):
            kmers = [kmers]
        uk = unique_kmers(kmers)
        if not uk:
            return {}
        klen = len(uk[0])
        n = self.index.num_cols
        packed = self.index.lookup_kmers(kmers_to_array(uk, klen), klen, self.num_hashes)
        nbits = n if remove_trailing_zeros else 8 * ((n + 7) // 8)
        return {km: _bits.from_packed(packed[i], nbits) for i, km in enumerate(uk)}

    def search(self, seq, threshold=1.0, score=False):
        """graph/bigsi.py:174-190.  Without scoring the whole filter stage (k-mer windows, set(kmers),
        min_kmers, hashing, row gather, AND / count, threshold) is ONE C-ABI call on the device
        (bigsi_b200_search_sequence); Python keeps the validation, the ordering rules and the result
        dictionaries."""
        assert threshold <= 1
        if not (isinstance(seq, str) and seq.isascii()):
            # The reference hashes the UTF-8 bytes of each k-character window (bloom/bloomfilter.py:5-6), so a window
            # with a non-ASCII character is longer than k bytes; this engine works on fixed k-byte windows.  DNA
            # queries are ASCII: anything else is rejected instead of being answered differently.
            raise ValueError("BIGSI.search takes an ASCII str sequence")
        n = self.num_samples
        colours, found, n_hits, num_kmers = self.index.search_sequence(seq.encode("ascii"), self.kmer_size, self.num_hashes,
                                                                       threshold, cap=max(self.index.num_cols, 1))
        if num_kmers <= self.min_unique_kmers_in_query:
            self.__warn_few_kmers(num_kmers)
        if num_kmers == 0:
            # the reference reduces over an empty list of per-k-mer vectors (utils/fncts.py:24-25)
            raise TypeError("reduce() of empty iterable with no initial value")
        keep = colours < n
        colours, found = colours[keep], found[keep]
        if threshold == 1.0:
            # exact_filter (graph/bigsi.py:192-205): ascending colour, every k-mer found
            names = self.colours_to_samples(colours.tolist())
            results = [BigsiQueryResult(colour=int(c), sample_name=names[int(c)], num_kmers=num_kmers, num_kmers_found=num_kmers)
                       for c in colours]
        else:
            # inexact_filter (graph/bigsi.py:211-230): stable sort by count, descending
            order = np.argsort(-found.astype(np.int64), kind="stable")
            results = [BigsiQueryResult(colour=int(colours[i]), sample_name=self.colour_to_sample(int(colours[i])),
                                        num_kmers_found=int(found[i]), num_kmers=num_kmers) for i in order]
        if score:
            self.score(seq, results)
        return [r.todict() for r in results if not r.sample_name == DELETION_SPECIAL_SAMPLE_NAME]

    def exact_filter(self, kmer_array, num_kmers):
        """graph/bigsi.py:192-205: colours whose column is set for every k-mer, ascending."""
        n = self.num_samples
        presence = self.index.search_kmers(kmer_array, self.kmer_size, self.num_hashes, mode=MODE_AND)[0]
        colours = np.nonzero(np.unpackbits(presence)[:n])[0].tolist()
        names = self.colours_to_samples(colours)
        return [BigsiQueryResult(colour=c, sample_name=names[c], num_kmers=num_kmers, num_kmers_found=num_kmers)
                for c in colours]

    def inexact_filter(self, kmer_array, num_kmers, min_kmers):
        """graph/bigsi.py:211-230: per-sample k-mer counts >= min_kmers, stable sort by count desc."""
        n = self.num_samples
        if min_kmers <= 0:
            counts = self.index.search_kmers(kmer_array, self.kmer_size, self.num_hashes, mode=MODE_COUNTS)[0][:n]
            colours, found = np.arange(n), counts
        else:
            colours, found, total = self.index.search_kmers_hits(kmer_array, self.kmer_size, self.num_hashes,
                                                                 [min(min_kmers, 0xFFFFFFFF)])[0]
            keep = colours < n
            colours, found = colours[keep], found[keep]
        order = np.argsort(-found.astype(np.int64), kind="stable")  # ties stay in ascending colour
        return [BigsiQueryResult(colour=int(colours[i]), sample_name=self.colour_to_sample(int(colours[i])),
                                 num_kmers_found=int(found[i]), num_kmers=num_kmers) for i in order]

    def score(self, seq, results):
        """graph/bigsi.py:232-239: for every hit, the presence of EVERY window of the query (in sequence
        order, duplicates included) in the hit's column -> Scorer.  The K x hits presence bits come from
        the GPU (bigsi_b200_sequence_presence) instead of a K x N int32 matrix built by repeated vstack."""
        if not results:
            return
        if len(seq) - self.kmer_size + 1 == 1:
            # a single window: the reference's unpack_and_cat yields a 1-D array and X[:, colour] raises
            raise IndexError("too many indices for array: array is 1-dimensional, but 2 were indexed")
        presence = self.index.sequence_presence(seq.encode("ascii"), self.kmer_size, self.num_hashes,
                                                [r.colour for r in results])
        for res, row in zip(results, presence):
            col = row.tobytes().decode("ascii")
            score_results = self.scorer.score(col)
            score_results["kmer-presence"] = col
            res.add_score(score_results)

    # -- persistence --------------------------------
    def _meta_entries(self):
        return [[key[1], value] for key, value in self._store.meta.items()]

    def save(self, path):
        """Flat index file (include/bigsi_b200.h "persistence"): rows in the reference's byte layout plus
        the ksi:* / metadata:* keys as JSON.  Replaces the durability of the reference's KV stores."""
        meta = {"format": "bigsi_b200/1", "k": self._store.kmer_size, "m": self.bloomfilter_size, "h": self.num_hashes,
                "metadata": self._meta_entries()}
        self.index.save(path, json.dumps(meta, separators=(",", ":")).encode("utf-8"))

    @classmethod
    def load(cls, config, path):
        """Open an index file written by save() into HBM (file -> pinned double buffer -> device) and
        register it under config's storage name.  config["k"/"m"/"h"] are taken from the file."""
        config = dict(config)  # k / m / h below come from the file: the caller's dict is not touched
        hd, meta_bytes = file_info(path)
        meta = json.loads(meta_bytes.decode("utf-8"))
        sc = config.get("storage-config", {}) or {}
        capacity = max(int(sc.get("col_capacity", 0)), hd["num_cols"], 1)
        index = make_index(hd["num_rows"], hd["num_cols"], col_capacity=capacity, col_offset=hd["col_offset"],
                           devices=_devices(config))
        try:
            index.load_rows(path, hd["rows_offset"], hd["row_bytes"], 0, 0, hd["num_rows"])
        except Exception:
            index.close()
            raise
        store = _Store(index, meta["m"], meta["h"], meta["k"])
        for key, value in meta["metadata"]:
            store.meta[("metadata", key)] = value
        config["k"], config["m"], config["h"] = meta["k"], meta["m"], meta["h"]
        _register(config, store)
        return cls(config)

    def to_kv(self, rows_per_chunk=4096):
        """The index as the reference's v0.3 key/value schema (storage/base.py:29-52,77-94; graph/index.py:10-11;
        matrix/bitmatrix.py:3-4; graph/metadata.py:8-10,111-112): a dict a reference storage backend could
        be filled from (`storage[k] = v`)."""
        kv = {}

        def put_int(key, value):
            kv[("%s:int" % key).encode("utf-8")] = str(int(value)).encode("utf-8")

        put_int("ksi:bloomfilter_size", self.bloomfilter_size)
        put_int("ksi:num_hashes", self.num_hashes)
        put_int("number_of_rows", self.bloomfilter_size)
        put_int("number_of_cols", self.index.num_cols)
        for (_, key), value in self._store.meta.items():
            if isinstance(value, int):
                put_int("metadata:%s" % key, value)
            else:
                kv[("metadata:%s:string" % key).encode("utf-8")] = str(value).encode("utf-8")
        m = self.bloomfilter_size
        for r0 in range(0, m, rows_per_chunk):
            rows = self.index.download_rows(r0, min(rows_per_chunk, m - r0))
            for i in range(rows.shape[0]):
                kv[b"%d:bitarray" % (r0 + i)] = rows[i].tobytes()
        return kv

    @classmethod
    def from_kv(cls, config, kv, rows_per_chunk=4096):
        """Import an index from the reference's v0.3 key/value schema (any mapping with bytes keys, e.g.
        a dump of its RocksDB/BerkeleyDB store) into HBM and register it under config's storage name."""
        def get_int(key):
            return int(bytes(kv[("%s:int" % key).encode("utf-8")]).decode("utf-8"))

        config = dict(config)  # m / h below come from the store: the caller's dict is not touched
        m, h = get_int("ksi:bloomfilter_size"), get_int("ksi:num_hashes")
        n_rows, n_cols = get_int("number_of_rows"), get_int("number_of_cols")
        if n_rows != m:
            raise ValueError("number_of_rows=%d differs from ksi:bloomfilter_size=%d" % (n_rows, m))
        sc = config.get("storage-config", {}) or {}
        capacity = max(int(sc.get("col_capacity", 0)), n_cols, 1)
        index = make_index(m, n_cols, col_capacity=capacity, devices=_devices(config))
        store = _Store(index, m, h, config["k"])
        try:
            row_bytes = (n_cols + 7) // 8
            for r0 in range(0, m, rows_per_chunk):
                n = min(rows_per_chunk, m - r0)
                buf = np.zeros((n, max(row_bytes, 1)), dtype=np.uint8)
                for i in range(n):
                    v = np.frombuffer(bytes(kv[b"%d:bitarray" % (r0 + i)]), dtype=np.uint8)
                    buf[i, : min(v.size, row_bytes)] = v[:row_bytes]
                if row_bytes:
                    index.upload_rows(r0, buf)
            for key, value in kv.items():
                key = bytes(key).decode("utf-8")
                if not key.startswith("metadata:"):
                    continue
                name, kind = key[len("metadata:"):].rsplit(":", 1)
                if kind == "int":  # sample name -> colour, or the colour count
                    store.meta[("metadata", name)] = int(bytes(value).decode("utf-8"))
                else:              # colour -> sample name
                    store.meta[("metadata", int(name))] = bytes(value).decode("utf-8")
        except Exception:
            store.close()
            raise
        config["m"], config["h"] = m, h
        _register(config, store)
        return cls(config)

    def __warn_few_kmers(self, n):
        logger.warning(
            "Query string should contain at least %i unique kmers. Your query contained %i unique kmers, and as a "
            "result the false discovery rate may be high. In future this will become an error."
            % (self.min_unique_kmers_in_query, n)
        )

    def __validate_search_query(self, seq):
        kmers = set()
        for k in self.seq_to_kmers(seq):
            kmers.add(k)
            if len(kmers) > self.min_unique_kmers_in_query:
                return True
        logger.warning(
            "Query string should contain at least %i unique kmers. Your query contained %i unique kmers, and as a "
            "result the false discovery rate may be high. In future this will become an error."
            % (self.min_unique_kmers_in_query, len(kmers))
        )
