"""ShardedIndex: the m x N signature matrix column-sharded over several GPUs of one box, driven by ONE process.

The reference keeps one row store per index (bigsi/storage/__init__.py:18-19) and has no distributed path
(SURVEY.md section 2.3).  Every sample column is independent in both the AND and the count
(bigsi/graph/index.py:42-80, graph/bigsi.py:192-230), so shard g holds ALL m rows of a contiguous range of
columns on its own GPU, global colour = col_offset + local column, and a search is the same search on every shard:

  * `search_sequence`: the sequence is submitted to every shard (bigsi_b200_search_sequence_submit: each GPU finds the
    unique windows itself inside its gather kernel, so the shards never talk to each other), then every shard's
    ticket is collected; per-shard hit lists are concatenated in shard order, which IS ascending colour order.
  * row-level calls (upload / download / save / load, the KV schema) cut or join the reference's row bytes at the
    shards' byte offsets (col_offset is always a multiple of 8).

The class mirrors the part of DeviceIndex that bigsi_b200.BIGSI uses, so `storage-config: {"devices": [0, 1, ...]}`
is the only thing a user changes.  (Multi-PROCESS deployments -- one rank per GPU under torchrun, the layout
bench.py measures -- use bigsi_b200.sharded.FusedExchange instead, where the query kernels exchange the query and
the hits over NVLink themselves.)
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import MODE_AND, MODE_COUNTS
from .index import DeviceIndex
from .sharded import shard_columns


class ShardedIndex:
    def __init__(self, num_rows, num_cols, devices, col_capacity=0, shards=None):
        """Empty m x num_cols matrix over len(devices) shards: balanced contiguous column ranges, every range start a
        multiple of 8.  col_capacity (optional) is the TOTAL number of columns the shards shall be able to hold without
        re-pitching; spare capacity goes to the last shard (appends land there)."""
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise ValueError("storage-config.devices must name at least one device")
        self.m = int(num_rows)
        if shards is not None:
            self.shards = shards
            return
        ranges = shard_columns(int(num_cols), len(self.devices))
        self.shards = []
        try:
            for g, (a, b) in enumerate(ranges):
                last = g == len(ranges) - 1
                cap = (b - a) + (max(int(col_capacity) - int(num_cols), 0) if last else 0)
                # (an empty range -- fewer columns than shards -- is re-created at its real offset by the first append)
                self.shards.append(DeviceIndex(self.m, b - a, col_capacity=max(cap, 1), col_offset=a if b > a else 0,
                                               device=self.devices[g]))
        except Exception:
            self.close()
            raise

    # -- bookkeeping -------------------------------------------------------------------------------
    def close(self):
        for s in getattr(self, "shards", []):
            s.close()
        self.shards = []

    def _infos(self):
        return [s.info() for s in self.shards]

    def _live(self):
        """[(shard, col_offset, num_cols)] of the shards that hold columns."""
        return [(s, i["col_offset"], i["num_cols"]) for s, i in zip(self.shards, self._infos()) if i["num_cols"]]

    @property
    def num_rows(self):
        return self.m

    @property
    def num_cols(self):
        live = self._live()
        return live[-1][1] + live[-1][2] if live else 0

    @property
    def row_bytes(self):
        return (self.num_cols + 7) // 8

    def info(self):
        infos = self._infos()
        n = self.num_cols
        cap = n
        if infos:
            live = [i for i in infos if i["num_cols"]] or infos[:1]
            cap = live[-1]["col_offset"] + live[-1]["col_capacity"]
        return {"num_rows": self.m, "num_cols": n, "col_capacity": cap, "col_offset": 0, "row_bytes": (n + 7) // 8,
                "device": self.devices[0], "devices": list(self.devices), "shards": infos,
                "matrix_bytes": sum(i["matrix_bytes"] for i in infos),
                "kernel_launches": sum(i["kernel_launches"] for i in infos)}

    def set_option(self, key, value):
        for s in self.shards:
            s.set_option(key, value)

    # -- matrix content ----------------------------------------------------------------------------
    def upload_rows(self, row0, rows, src_byte_offset=0):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        for s, off, n in self._live():
            s.upload_rows(row0, rows, src_byte_offset=src_byte_offset + off // 8)

    def download_rows(self, row0, n_rows):
        out = np.zeros((n_rows, self.row_bytes), dtype=np.uint8)
        for s, off, n in self._live():
            out[:, off // 8 : off // 8 + (n + 7) // 8] = s.download_rows(row0, n_rows)
        return out

    def build_columns(self, col0, blooms, n_bits=None):
        """Bloom filters -> GLOBAL columns [col0, col0 + n) (BIGSI.build / bulk insert), each shard transposing its own."""
        blooms = np.ascontiguousarray(blooms, dtype=np.uint8)
        c1 = col0 + blooms.shape[0]
        for s, a, n in self._live():  # (the shards were created with their planned column counts)
            lo, hi = max(col0, a), min(c1, a + n)
            if lo < hi:
                s.build_columns(lo - a, blooms[lo - col0 : hi - col0], n_bits)

    def _append_target(self, col):
        """(shard, local column) where GLOBAL column `col` goes; col == num_cols appends."""
        infos = self._infos()
        for s, i in zip(self.shards, infos):
            if i["num_cols"] and i["col_offset"] <= col < i["col_offset"] + i["num_cols"]:
                return s, col - i["col_offset"]
        n = self.num_cols
        if col != n:
            raise IndexError("column %d beyond num_cols=%d" % (col, n))
        live = [g for g, i in enumerate(infos) if i["num_cols"]]
        g = live[-1] if live else 0
        i = infos[g]
        if i["num_cols"] < i["col_capacity"]:
            return self.shards[g], i["num_cols"]
        if g + 1 < len(self.shards):  # the next shard starts exactly where this one is full (a multiple of 8)
            nxt = self.shards[g + 1]
            ni = infos[g + 1]
            self.shards[g + 1] = DeviceIndex(self.m, 0, col_capacity=max(ni["col_capacity"], 1), col_offset=n, device=self.devices[g + 1])
            nxt.close()
            return self.shards[g + 1], 0
        self.shards[g] = self.shards[g].grown(2 * i["col_capacity"])
        return self.shards[g], i["num_cols"]

    def set_column(self, col, bloom_packed, n_bits):
        s, local = self._append_target(col)
        s.set_column(local, bloom_packed, n_bits)

    def grown(self, new_capacity):
        """Capacity is managed per shard (set_column grows the last shard when it has to)."""
        return self

    # -- queries -----------------------------------------------------------------------------------
    def search_sequence(self, seq, k, h, threshold, cap=None):
        live = self._live() or [(self.shards[0], 0, 0)]  # (an index without samples still counts the query's k-mers)
        tickets = [(s, off, s.search_sequence_submit(seq, k, h, threshold, cap=n if cap is None else min(int(cap), n)))
                   for s, off, n in live]
        cols, cnts, n_hits, U = [], [], 0, 0
        for s, off, t in tickets:
            c, v, nh, u = s.search_sequence_wait(t)
            cols.append(c.astype(np.int64) + off)
            cnts.append(v)
            n_hits += nh
            U = u
        return np.concatenate(cols), np.concatenate(cnts), n_hits, U

    def search_kmers(self, kmers, k, h, q_offsets=None, mode=MODE_COUNTS):
        parts = [(off, n, s.search_kmers(kmers, k, h, q_offsets, mode)) for s, off, n in self._live()]
        nq = parts[0][2].shape[0] if parts else (1 if q_offsets is None else len(q_offsets) - 1)
        if mode == MODE_COUNTS:
            out = np.zeros((nq, self.num_cols), dtype=np.uint32)
            for off, n, p in parts:
                out[:, off : off + n] = p
            return out
        out = np.zeros((nq, self.row_bytes), dtype=np.uint8)
        for off, n, p in parts:
            out[:, off // 8 : off // 8 + (n + 7) // 8] = p
        return out

    def search_kmers_hits(self, kmers, k, h, min_kmers, q_offsets=None, cap=None):
        per = [(off, s.search_kmers_hits(kmers, k, h, min_kmers, q_offsets, cap)) for s, off, n in self._live()]
        nq = len(per[0][1]) if per else 0
        out = []
        for q in range(nq):
            cols = np.concatenate([r[q][0].astype(np.int64) + off for off, r in per])
            cnts = np.concatenate([r[q][1] for off, r in per])
            out.append((cols, cnts, sum(r[q][2] for off, r in per)))
        return out

    def lookup_kmers(self, kmers, k, h):
        live = self._live()
        parts = [(off, n, s.lookup_kmers(kmers, k, h)) for s, off, n in live]
        n_kmers = parts[0][2].shape[0] if parts else len(kmers)
        out = np.zeros((n_kmers, self.row_bytes), dtype=np.uint8)
        for off, n, p in parts:
            out[:, off // 8 : off // 8 + (n + 7) // 8] = p
        return out

    def sequence_presence(self, seq, k, h, cols):
        cols = np.asarray(cols, dtype=np.int64)
        arr = DeviceIndex._seq_array(seq)
        out = np.zeros((cols.size, max(arr.size - k + 1, 0)), dtype=np.uint8)
        for s, off, n in self._live():
            sel = np.nonzero((cols >= off) & (cols < off + n))[0]
            if sel.size:
                out[sel] = s.sequence_presence(seq, k, h, (cols[sel] - off).astype(np.int32))
        return out

    # -- persistence: ONE full-width file, the same format a single-GPU index writes ------------------
    def save(self, path, meta=b"", rows_per_chunk=None):
        """Header + metadata + m full-width rows (include/bigsi_b200.h "persistence"): the shards' byte ranges are
        joined on the host chunk by chunk, so a file written from 8 GPUs loads onto 1 and vice versa."""
        meta = bytes(meta)
        n, rb = self.num_cols, self.row_bytes
        hd = _lib.FileHeader()
        hd.magic = b"BIGSIB2\n"
        hd.version, hd.header_bytes = 1, ctypes.sizeof(_lib.FileHeader)
        hd.num_rows, hd.num_cols, hd.col_offset, hd.row_bytes, hd.meta_bytes = self.m, n, 0, rb, len(meta)
        hd.rows_offset = -(-(ctypes.sizeof(_lib.FileHeader) + len(meta)) // 4096) * 4096
        step = rows_per_chunk or max(1, (64 << 20) // max(rb, 1))
        with open(path, "wb") as f:
            f.write(bytes(hd))
            f.write(meta)
            f.write(b"\0" * (hd.rows_offset - ctypes.sizeof(_lib.FileHeader) - len(meta)))
            if rb:
                for r0 in range(0, self.m, step):
                    f.write(self.download_rows(r0, min(step, self.m - r0)).tobytes())

    def load_rows(self, path, file_offset, file_stride, src_byte_offset=0, row0=0, n_rows=None):
        """Every shard reads its own byte range of the file's full-width rows (native double-buffered upload)."""
        for s, off, n in self._live():
            s.load_rows(path, file_offset, file_stride, src_byte_offset + off // 8, row0, n_rows)


def make_index(num_rows, num_cols, col_capacity=0, col_offset=0, device=0, devices=None):
    """DeviceIndex on one GPU, or a ShardedIndex when `devices` names several."""
    if devices is not None and len(devices) > 1:
        if col_offset:
            raise ValueError("a sharded index starts at colour 0")
        return ShardedIndex(num_rows, num_cols, devices, col_capacity=col_capacity)
    if devices:
        device = devices[0]
    return DeviceIndex(num_rows, num_cols, col_capacity=col_capacity, col_offset=col_offset, device=device)
