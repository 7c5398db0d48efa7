"""DeviceIndex: one column shard of the m x N signature matrix resident in HBM.

Python face of the C ABI (include/bigsi_b200.h).  Replaces the reference's storage backends
(bigsi/storage/*.py) plus BitMatrix (bigsi/matrix/bitmatrix.py:7-75) for the search path: rows are
addressed by number in one packed HBM array instead of by key in a KV store.
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import MODE_AND, MODE_COUNTS, check


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None and a.size else ctypes.c_void_p(0)


def kmers_to_array(kmers, k):
    """list of k-mer strings (all of length k, ASCII) -> uint8 [n, k]."""
    n = len(kmers)
    if n == 0:
        return np.zeros((0, k), dtype=np.uint8)
    joined = "".join(kmers).encode("utf-8")
    if len(joined) != n * k:
        raise ValueError("k-mers must be ASCII strings of length k=%d" % k)
    return np.frombuffer(joined, dtype=np.uint8).reshape(n, k)


def hash_kmers(kmers, k, h, m, canonical=True, device=0):
    """Row ids int32 [n, h] of n k-mers (bigsi/bloom/bloomfilter.py:5-13 after
    utils/fncts.py:47-54 when canonical) computed by the CUDA hash kernel."""
    arr = kmers if isinstance(kmers, np.ndarray) else kmers_to_array(list(kmers), k)
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    out = np.empty((arr.shape[0], h), dtype=np.int32)
    check(_lib.lib().bigsi_b200_hash_kmers(device, _ptr(arr), arr.shape[0], k, h, m, 1 if canonical else 0, _ptr(out)))
    return out


def bloom_kmers(kmers, k, h, m, canonical=True, device=0):
    """Packed MSB-first Bloom filter bytes (uint8 [ceil(m/8)]) of n k-mers, built on the GPU
    (bigsi/bloom/bloomfilter.py:16-32; .bloom file layout of bigsi/cmds/bloom.py:26-27)."""
    arr = kmers if isinstance(kmers, np.ndarray) else kmers_to_array(list(kmers), k)
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    out = np.zeros((m + 7) // 8, dtype=np.uint8)
    check(_lib.lib().bigsi_b200_bloom_kmers(device, _ptr(arr), arr.shape[0] if arr.ndim == 2 else 0, k, h, m,
                                            1 if canonical else 0, _ptr(out)))
    return out


def file_info(path):
    """(header dict, metadata bytes) of an index file written by DeviceIndex.save."""
    hd = _lib.FileHeader()
    L = _lib.lib()
    check(L.bigsi_b200_file_info(os.fsencode(path), ctypes.byref(hd), None, 0))
    meta = ctypes.create_string_buffer(max(int(hd.meta_bytes), 1))
    check(L.bigsi_b200_file_info(os.fsencode(path), ctypes.byref(hd), meta, int(hd.meta_bytes)))
    d = {name: getattr(hd, name) for name, _ in hd._fields_ if name != "magic"}
    return d, meta.raw[: int(hd.meta_bytes)]


class DeviceIndex:
    def __init__(self, num_rows, num_cols, col_capacity=0, col_offset=0, device=0):
        self._h = ctypes.c_void_p(0)
        L = _lib.lib()
        check(L.bigsi_b200_index_create(device, num_rows, num_cols, col_capacity, col_offset, ctypes.byref(self._h)))
        self._L = L
        self._hit_bufs = {}
        self._one_query_offsets = np.zeros(2, dtype=np.int64)
        self._num_kmers_out = np.zeros(1, dtype=np.uint64)

    # -- lifecycle -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.bigsi_b200_index_destroy(self._h)
            self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h.value:
            raise ValueError("index has been destroyed")
        return self._h

    def info(self):
        i = _lib.Info()
        check(self._L.bigsi_b200_index_get_info(self.handle, ctypes.byref(i)))
        return i.asdict()

    @property
    def num_rows(self):
        return self.info()["num_rows"]

    @property
    def num_cols(self):
        return self.info()["num_cols"]

    @property
    def row_bytes(self):
        return self.info()["row_bytes"]

    def set_option(self, key, value):
        check(self._L.bigsi_b200_index_set_option(self.handle, key.encode(), int(value)))

    def timing_collect(self):
        """(fused_kernel_ms_sum, merge_kernel_ms_sum, n_launches) since the last collect."""
        a, b, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_uint64(0)
        check(self._L.bigsi_b200_index_timing_collect(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)))
        return a.value, b.value, n.value

    # -- matrix content ------------------------------------------------------
    def upload_rows(self, row0, rows, src_byte_offset=0):
        """rows: uint8 [n, >= row_bytes] in the reference's bitarray.tobytes() layout."""
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        if rows.ndim != 2:
            raise ValueError("rows must be 2-D (n_rows, bytes)")
        check(self._L.bigsi_b200_index_upload_rows(self.handle, row0, rows.shape[0], _ptr(rows), rows.strides[0],
                                                   src_byte_offset))

    def download_rows(self, row0, n_rows):
        out = np.empty((n_rows, self.row_bytes), dtype=np.uint8)
        check(self._L.bigsi_b200_index_download_rows(self.handle, row0, n_rows, _ptr(out), out.strides[0] if n_rows else 0))
        return out

    def set_column(self, col, bloom_packed, n_bits):
        bloom_packed = np.ascontiguousarray(bloom_packed, dtype=np.uint8)
        if bloom_packed.size * 8 < n_bits:
            raise ValueError("bloom filter shorter than n_bits")
        check(self._L.bigsi_b200_index_set_column(self.handle, col, _ptr(bloom_packed), n_bits))

    def build_columns(self, col0, blooms, n_bits=None):
        """BIGSI.build's transpose / bulk insert on the device: blooms = uint8 [n, >= ceil(n_bits/8)]
        packed MSB-first Bloom filters -> local columns [col0, col0 + n)."""
        blooms = np.ascontiguousarray(blooms, dtype=np.uint8)
        if blooms.ndim != 2:
            raise ValueError("blooms must be 2-D (n_filters, bytes)")
        n_bits = self.num_rows if n_bits is None else int(n_bits)
        if blooms.shape[1] * 8 < n_bits:
            raise ValueError("bloom filters shorter than n_bits")
        check(self._L.bigsi_b200_index_build_columns(self.handle, col0, blooms.shape[0], _ptr(blooms),
                                                     blooms.strides[0] if blooms.shape[0] else 0, n_bits))

    def build_columns_dev(self, col0, n_blooms, d_blooms, bloom_stride, n_bits, stream=0):
        check(self._L.bigsi_b200_index_build_columns_dev(self.handle, col0, n_blooms, d_blooms, bloom_stride, n_bits, stream))

    def grown(self, new_capacity):
        """A re-pitched copy of this shard that holds `new_capacity` columns (host round trip in chunks); closes self.
        BitMatrix.insert_column appends to rows of any length (matrix/bitmatrix.py:67-75, storage/base.py:113-116);
        here the pitch is fixed at creation, so an insert beyond it moves the matrix once (capacity doubles)."""
        info = self.info()
        new = DeviceIndex(info["num_rows"], info["num_cols"], col_capacity=new_capacity, col_offset=info["col_offset"],
                          device=info["device"])
        try:
            step = max(1, (1 << 26) // max(info["row_bytes"], 1))
            for r0 in range(0, info["num_rows"], step):
                n = min(step, info["num_rows"] - r0)
                new.upload_rows(r0, self.download_rows(r0, n))
        except Exception:
            new.close()
            raise
        self.close()
        return new

    def save(self, path, meta=b""):
        """Write the shard to a flat index file (include/bigsi_b200.h "persistence")."""
        meta = bytes(meta)
        check(self._L.bigsi_b200_index_save(self.handle, os.fsencode(path), meta, len(meta)))

    def load_rows(self, path, file_offset, file_stride, src_byte_offset=0, row0=0, n_rows=None):
        n_rows = self.num_rows - row0 if n_rows is None else n_rows
        check(self._L.bigsi_b200_index_load_rows(self.handle, os.fsencode(path), file_offset, file_stride, src_byte_offset,
                                                 row0, n_rows))

    def sequence_presence(self, seq, k, h, cols):
        """score=True support: uint8 [len(cols), len(seq)-k+1] of the characters '0'/'1' -- row c is the
        reference's "kmer-presence" string of local column cols[c] (graph/bigsi.py:232-239)."""
        arr = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        n = max(arr.size - k + 1, 0)
        out = np.zeros((cols.size, n), dtype=np.uint8)
        if n and cols.size:
            check(self._L.bigsi_b200_sequence_presence(self.handle, _ptr(arr), arr.size, k, h, _ptr(cols), cols.size, _ptr(out)))
        return out

    def fill_synthetic(self, seed=0, and_draws=1, planted_cols=(), planted_thr=()):
        pc = np.ascontiguousarray(planted_cols, dtype=np.uint64)
        pt = np.ascontiguousarray(planted_thr, dtype=np.uint32)
        if pc.shape != pt.shape:
            raise ValueError("planted_cols and planted_thr must have the same length")
        check(self._L.bigsi_b200_index_fill_synthetic(self.handle, seed, and_draws, _ptr(pc), _ptr(pt), pc.size))

    # -- queries (host buffers) ----------------------------------------------
    @staticmethod
    def _offsets(q_offsets, n):
        if q_offsets is None:
            q_offsets = [0, n]
        q = np.ascontiguousarray(q_offsets, dtype=np.int64)
        if q.ndim != 1 or q.size < 1 or q[0] != 0 or q[-1] != n:
            raise ValueError("q_offsets must start at 0 and end at the number of k-mers")
        return q

    def _out(self, mode, nq):
        nc = self.num_cols
        if mode == MODE_COUNTS:
            return np.zeros((nq, max(nc, 1)), dtype=np.uint32), max(nc, 1)
        rb = max((nc + 7) // 8, 1)
        return np.zeros((nq, rb), dtype=np.uint8), rb

    def search_kmers(self, kmers, k, h, q_offsets=None, mode=MODE_COUNTS):
        """Unique raw k-mers of a batch of queries -> uint32 counts [Q, N] or packed MSB-first
        presence bytes [Q, ceil(N/8)]."""
        arr = kmers if isinstance(kmers, np.ndarray) else kmers_to_array(list(kmers), k)
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        q = self._offsets(q_offsets, arr.shape[0])
        out, stride = self._out(mode, q.size - 1)
        check(self._L.bigsi_b200_search_kmers(self.handle, mode, _ptr(arr), _ptr(q), q.size - 1, k, h, _ptr(out), stride))
        nc = self.num_cols
        return out[:, :nc] if mode == MODE_COUNTS else out[:, : (nc + 7) // 8]

    def search_rows(self, rows, h, q_offsets=None, mode=MODE_COUNTS):
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, h)
        q = self._offsets(q_offsets, rows.shape[0])
        out, stride = self._out(mode, q.size - 1)
        check(self._L.bigsi_b200_search_rows(self.handle, mode, _ptr(rows), _ptr(q), q.size - 1, h, _ptr(out), stride))
        nc = self.num_cols
        return out[:, :nc] if mode == MODE_COUNTS else out[:, : (nc + 7) // 8]

    def search_kmers_hits(self, kmers, k, h, min_kmers, q_offsets=None, cap=None):
        """Fused search + threshold.  Returns a list (one per query) of (colours int32, counts
        uint32, n_hits) with count >= min_kmers[q], colours ascending."""
        arr = kmers if isinstance(kmers, np.ndarray) else kmers_to_array(list(kmers), k)
        if arr.dtype != np.uint8 or not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr, dtype=np.uint8)
        if q_offsets is None:
            nq = 1
            q = self._one_query_offsets
            q[1] = arr.shape[0]
        else:
            q = self._offsets(q_offsets, arr.shape[0])
            nq = q.size - 1
        mk = np.ascontiguousarray(np.broadcast_to(np.asarray(min_kmers, dtype=np.uint32), (nq,)))
        cap = self.num_cols if cap is None else int(cap)
        key = (nq, cap)
        bufs = self._hit_bufs.get(key)
        if bufs is None:  # output staging is reused between calls (results are copied out below)
            bufs = (np.empty((nq, max(cap, 1)), dtype=np.int32), np.empty((nq, max(cap, 1)), dtype=np.uint32),
                    np.zeros(nq, dtype=np.uint64))
            if len(self._hit_bufs) > 8:
                self._hit_bufs.clear()
            self._hit_bufs[key] = bufs
        cols, cnts, n = bufs
        check(self._L.bigsi_b200_search_kmers_hits(self.handle, arr.ctypes.data, q.ctypes.data, nq, k, h, mk.ctypes.data,
                                                   cols.ctypes.data, cnts.ctypes.data, cap, n.ctypes.data))
        res = []
        for i in range(nq):
            ni = int(n[i])
            m = min(ni, cap)
            c, v = cols[i, :m], cnts[i, :m]
            order = np.argsort(c, kind="stable")
            res.append((c[order], v[order], ni))
        return res

    def search_sequence(self, seq, k, h, threshold, cap=None):
        """BIGSI.search's filter stage for one sequence, entirely on the device (windows -> set of raw
        k-mers -> min_kmers = ceil(U * threshold) -> hash -> gather-AND-count -> threshold).
        seq: bytes / bytearray / uint8 array.  Returns (colours int32 ascending, counts uint32,
        n_hits, num_kmers U); U == 0 when the sequence is shorter than k."""
        if isinstance(seq, (bytes, bytearray)):
            arr = np.frombuffer(seq, dtype=np.uint8)
        else:
            arr = np.ascontiguousarray(seq, dtype=np.uint8)
        cap = self.num_cols if cap is None else int(cap)
        key = (1, cap)
        bufs = self._hit_bufs.get(key)
        if bufs is None:
            bufs = (np.empty((1, max(cap, 1)), dtype=np.int32), np.empty((1, max(cap, 1)), dtype=np.uint32),
                    np.zeros(1, dtype=np.uint64))
            if len(self._hit_bufs) > 8:
                self._hit_bufs.clear()
            self._hit_bufs[key] = bufs
        cols, cnts, n = bufs
        u = self._num_kmers_out
        check(self._L.bigsi_b200_search_sequence(self.handle, arr.ctypes.data if arr.size else 0, arr.size, k, h,
                                                 float(threshold), cols.ctypes.data, cnts.ctypes.data, cap, n.ctypes.data,
                                                 u.ctypes.data))
        ni = int(n[0])
        m = min(ni, cap)
        c, v = cols[0, :m], cnts[0, :m]
        order = np.argsort(c, kind="stable")
        return c[order], v[order], ni, int(u[0])

    @staticmethod
    def _seq_array(seq):
        if isinstance(seq, (bytes, bytearray)):
            return np.frombuffer(seq, dtype=np.uint8)
        return np.ascontiguousarray(seq, dtype=np.uint8)

    def search_sequence_submit(self, seq, k, h, threshold, cap=None):
        """First half of search_sequence: stage the sequence and launch; returns a ticket for search_sequence_wait.
        Up to 8 tickets may be outstanding; one host thread can so keep several handles (column shards) busy."""
        arr = self._seq_array(seq)
        cap = self.num_cols if cap is None else int(cap)
        ticket = ctypes.c_uint64(0)
        check(self._L.bigsi_b200_search_sequence_submit(self.handle, arr.ctypes.data if arr.size else 0, arr.size, k, h,
                                                        float(threshold), cap, ctypes.byref(ticket)))
        return ticket.value, cap

    def search_sequence_wait(self, ticket):
        """(colours int32 ascending, counts uint32, n_hits, num_kmers U) of a submitted search."""
        ticket, cap = ticket
        cols = np.empty(max(cap, 1), dtype=np.int32)
        cnts = np.empty(max(cap, 1), dtype=np.uint32)
        n, u = ctypes.c_uint64(0), ctypes.c_uint64(0)
        check(self._L.bigsi_b200_search_sequence_wait(self.handle, ticket, cols.ctypes.data, cnts.ctypes.data, cap,
                                                      ctypes.byref(n), ctypes.byref(u)))
        m = min(n.value, cap)
        order = np.argsort(cols[:m], kind="stable")
        return cols[:m][order], cnts[:m][order], n.value, u.value

    def search_sequences(self, seqs, k, h, threshold, cap=1024):
        """bulk_search: a list of sequences (bytes) through one C-ABI call, pipelined on the device.  Returns a list of
        (colours int32 ascending, counts uint32, n_hits, num_kmers U), one per sequence; hit lists longer than `cap`
        are cut (n_hits stays exact)."""
        nq = len(seqs)
        if nq == 0:
            return []
        offsets = np.zeros(nq + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
        blob = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8) if offsets[-1] else np.zeros(1, dtype=np.uint8)
        cap = int(cap)
        cols = np.empty((nq, max(cap, 1)), dtype=np.int32)
        cnts = np.empty((nq, max(cap, 1)), dtype=np.uint32)
        n = np.zeros(nq, dtype=np.uint64)
        u = np.zeros(nq, dtype=np.uint64)
        check(self._L.bigsi_b200_search_sequences(self.handle, blob.ctypes.data, offsets.ctypes.data, nq, k, h, float(threshold),
                                                  cols.ctypes.data, cnts.ctypes.data, cap, n.ctypes.data, u.ctypes.data))
        out = []
        for q in range(nq):
            m = min(int(n[q]), cap)
            order = np.argsort(cols[q, :m], kind="stable")
            out.append((cols[q, :m][order], cnts[q, :m][order], int(n[q]), int(u[q])))
        return out

    def lookup_kmers(self, kmers, k, h):
        """Per-k-mer AND vectors: uint8 [n, ceil(N/8)] (graph/index.py:42-49)."""
        arr = kmers if isinstance(kmers, np.ndarray) else kmers_to_array(list(kmers), k)
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        rb = max(self.row_bytes, 1)
        out = np.zeros((arr.shape[0], rb), dtype=np.uint8)
        check(self._L.bigsi_b200_lookup_kmers(self.handle, _ptr(arr), arr.shape[0], k, h, _ptr(out), rb))
        return out[:, : self.row_bytes]

    # -- queries (device pointers; torch tensors or raw addresses) --------------
    def query_dev(self, mode, d_rows, d_q_offsets, n_queries, total_kmers, h, d_out, out_stride, stream=0,
                  max_query_kmers=0):
        check(self._L.bigsi_b200_query_dev(self.handle, mode, d_rows, d_q_offsets, n_queries, total_kmers,
                                           max_query_kmers, h, d_out, out_stride, stream))

    def query_hits_dev(self, d_rows, d_q_offsets, n_queries, total_kmers, h, d_min_kmers, d_cols_out, d_counts_out,
                       cap, d_n_out, stream=0, max_query_kmers=0, d_counts_full=0, counts_stride=0):
        check(self._L.bigsi_b200_query_hits_dev(self.handle, d_rows, d_q_offsets, n_queries, total_kmers,
                                                max_query_kmers, h, d_min_kmers, d_cols_out, d_counts_out, cap,
                                                d_n_out, d_counts_full, counts_stride, stream))

    def query_kmers_hits_dev(self, d_kmers, k, d_q_offsets, n_queries, total_kmers, h, d_min_kmers, d_cols_out,
                             d_counts_out, cap, d_n_out, stream=0, max_query_kmers=0, d_counts_full=0, counts_stride=0):
        check(self._L.bigsi_b200_query_kmers_hits_dev(self.handle, d_kmers, k, d_q_offsets, n_queries, total_kmers,
                                                      max_query_kmers, h, d_min_kmers, d_cols_out, d_counts_out, cap,
                                                      d_n_out, d_counts_full, counts_stride, stream))

    def query_kmers_hits_stream_dev(self, d_kmers, k, n_kmers, h, min_kmers, d_cols_out, d_counts_out, cap, d_n_out, stream=0):
        """ONE query, DEFERRED (include/bigsi_b200.h "streamed single-query launches"): the result is complete in stream
        order after the next streamed query on this handle or after flush()."""
        check(self._L.bigsi_b200_query_kmers_hits_stream_dev(self.handle, d_kmers, k, n_kmers, h, int(min_kmers), d_cols_out,
                                                             d_counts_out, cap, d_n_out, stream))

    def flush(self):
        """Launch stage 2 of the pending deferred query, if any (no host synchronisation)."""
        check(self._L.bigsi_b200_index_flush(self.handle))

    def lookup_dev(self, d_rows, n_kmers, h, d_out, out_stride, stream=0):
        check(self._L.bigsi_b200_lookup_dev(self.handle, d_rows, n_kmers, h, d_out, out_stride, stream))


def hash_kmers_dev(d_kmers, n, k, h, m, d_rows_out, stream=0, canonical=True):
    check(_lib.lib().bigsi_b200_hash_kmers_dev(d_kmers, n, k, h, m, 1 if canonical else 0, d_rows_out, stream))


def threshold_dev(d_counts, counts_stride, n_queries, num_cols, d_min_kmers, d_cols_out, d_counts_out, cap, d_n_out,
                  stream=0):
    check(_lib.lib().bigsi_b200_threshold_dev(d_counts, counts_stride, n_queries, num_cols, d_min_kmers, d_cols_out,
                                              d_counts_out, cap, d_n_out, stream))
