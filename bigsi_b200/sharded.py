"""Column (sample) sharding across the GPUs of one box: one process per GPU.

The reference has no distributed path; every sample column is independent
in both the AND and the count (bigsi/graph/index.py:42-80, graph/bigsi.py:192-230), so rank g
holds all m rows of the columns [g*N/G, (g+1)*N/G) and a query needs exactly two small exchanges:

  1. broadcast of the query (raw k-mer bytes uint8 [U*k]; the row ids int32 [U*h] are derived from
     them identically on every rank, inside the query kernel) from rank 0,
  2. all-gather of the per-shard hits (count + compact (colour, count) pairs) or counts.

torch.distributed is the plumbing (NCCL on the GPU box, gloo in the CPU tests); the local shard
object does the compute (DeviceShard = the CUDA path; tests inject an oracle-backed stand-in to
check the exchange logic on CPU).
"""
import numpy as np


def shard_columns(num_cols, world_size, align=8):
    """Contiguous column ranges [(start, stop)] per rank; every start is a multiple of `align`
    (shard col_offset must be byte aligned in the MSB-first row layout)."""
    per = -(-num_cols // world_size)
    per = -(-per // align) * align
    out = []
    for g in range(world_size):
        a = min(g * per, num_cols)
        b = min(a + per, num_cols)
        out.append((a, b))
    return out


def merge_shard_hits(n_per_rank, cols_per_rank, counts_per_rank, col_offsets):
    """Concatenate the per-shard hit lists into global colours, ascending (the order
    graph/bigsi.py:211-230 sorts from)."""
    cols, cnts = [], []
    for n, c, v, off in zip(n_per_rank, cols_per_rank, counts_per_rank, col_offsets):
        n = int(n)
        cols.append(np.asarray(c[:n], dtype=np.int64) + int(off))
        cnts.append(np.asarray(v[:n], dtype=np.int64))
    cols = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    cnts = np.concatenate(cnts) if cnts else np.zeros(0, dtype=np.int64)
    order = np.argsort(cols, kind="stable")
    return cols[order], cnts[order]


class DeviceShard:
    """The CUDA compute of one rank on torch tensors (device pointers go straight to the C ABI)."""

    def __init__(self, index, k, h, cap=1024):
        import torch

        self.torch = torch
        self.index = index
        info = index.info()
        self.k, self.h = k, h
        self.m = info["num_rows"]
        self.num_cols = info["num_cols"]
        self.col_offset = info["col_offset"]
        self.device = torch.device("cuda", info["device"])
        self.cap = cap
        self._counts = None

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def hash(self, kmers_u8):
        """uint8 [U, k] device tensor -> int32 [U, h] row ids."""
        from .index import hash_kmers_dev

        U = kmers_u8.shape[0]
        rows = self.torch.empty((U, self.h), dtype=self.torch.int32, device=self.device)
        hash_kmers_dev(kmers_u8.data_ptr(), U, self.k, self.h, self.m, rows.data_ptr(), self._stream())
        return rows

    def counts(self, rows, q_offsets, n_queries, max_query_kmers=0):
        """int32 [U, h] rows + int64 [Q+1] offsets (device) -> uint32-as-int32 counts [Q, stride]."""
        t = self.torch
        stride = max((self.num_cols + 3) // 4 * 4, 4)
        if self._counts is None or self._counts.shape[0] < n_queries or self._counts.shape[1] != stride:
            self._counts = t.empty((n_queries, stride), dtype=t.int32, device=self.device)
        out = self._counts[:n_queries]
        self.index.query_dev(0, rows.data_ptr(), q_offsets.data_ptr(), n_queries, rows.shape[0], self.h,
                             out.data_ptr(), stride, self._stream(), max_query_kmers)
        return out

    def hits(self, counts, min_kmers):
        """counts [Q, stride], min_kmers int32 [Q] (device) -> ONE packed int32 device buffer
        [Q*2 (hit count, little-endian int64) | Q*cap colours | Q*cap counts], so that the
        exchange is a single all-gather.  Order within a query is unspecified."""
        from .index import threshold_dev

        t = self.torch
        Q = counts.shape[0]
        buf = t.empty((Q * (2 + 2 * self.cap),), dtype=t.int32, device=self.device)
        base = buf.data_ptr()
        threshold_dev(counts.data_ptr(), counts.shape[1], Q, self.num_cols, min_kmers.data_ptr(),
                      base + 8 * Q, base + 8 * Q + 4 * Q * self.cap, self.cap, base, self._stream())
        return buf


    def search_hits(self, rows, q_offsets, n_queries, min_kmers, max_query_kmers=0):
        """Fused gather-AND-count + threshold (one C-ABI call, two kernels): same packed layout as
        hits(); the full count vectors are never written."""
        t = self.torch
        buf = t.empty((n_queries * (2 + 2 * self.cap),), dtype=t.int32, device=self.device)
        base = buf.data_ptr()
        self.index.query_hits_dev(rows.data_ptr(), q_offsets.data_ptr(), n_queries, rows.shape[0], self.h,
                                  min_kmers.data_ptr(), base + 8 * n_queries,
                                  base + 8 * n_queries + 4 * n_queries * self.cap, self.cap, base, self._stream(),
                                  max_query_kmers)
        return buf


    def search_kmers_hits(self, kmers_u8, q_offsets, n_queries, min_kmers, max_query_kmers=0):
        """The whole path in (normally) one kernel: raw k-mers uint8 [U, k] on the device ->
        canonical + murmur3 in the kernel prologue -> gather-AND-count -> grid barrier -> merge +
        threshold.  Same packed layout as hits()."""
        t = self.torch
        buf = t.empty((n_queries * (2 + 2 * self.cap),), dtype=t.int32, device=self.device)
        base = buf.data_ptr()
        self.index.query_kmers_hits_dev(kmers_u8.data_ptr(), self.k, q_offsets.data_ptr(), n_queries, kmers_u8.shape[0],
                                        self.h, min_kmers.data_ptr(), base + 8 * n_queries,
                                        base + 8 * n_queries + 4 * n_queries * self.cap, self.cap, base, self._stream(),
                                        max_query_kmers)
        return buf


def unpack_hits(buf, n_queries, cap):
    """Inverse of DeviceShard.hits' packing for a [G, Q*(2+2*cap)] (or 1-D) int32 array on the host."""
    a = np.asarray(buf).reshape(-1, n_queries * (2 + 2 * cap))
    n = np.ascontiguousarray(a[:, : 2 * n_queries]).view(np.int64).reshape(-1, n_queries)
    cols = a[:, 2 * n_queries : 2 * n_queries + n_queries * cap].reshape(-1, n_queries, cap)
    vals = a[:, 2 * n_queries + n_queries * cap :].reshape(-1, n_queries, cap)
    return n, cols, vals


class ShardedSearcher:
    """One rank's view of a column-sharded search.  `shard` provides hash/counts/hits on tensors
    living on `shard.device`; `dist` is torch.distributed (already initialised) or None for a
    single process."""

    def __init__(self, shard, dist=None, world_size=1, rank=0):
        self.shard = shard
        self.dist = dist if world_size > 1 else None
        self.world_size = world_size
        self.rank = rank
        self.torch = shard.torch

    def search_step(self, kmers_u8, q_offsets, min_kmers, n_queries, max_query_kmers=0):
        """One batched search: rank 0's k-mers decide; returns the packed hit buffers of all ranks,
        int32 [G, Q*(2+2*cap)] on the device (see DeviceShard.hits / unpack_hits; LOCAL colours)."""
        t = self.torch
        sh = self.shard
        if self.dist is not None:
            # exchange 1: the query itself (raw k-mer bytes; every rank derives the same row ids from
            # them inside its own kernel, so the ranks stay symmetric and rank 0 runs no extra kernel)
            if self.rank != 0:
                kmers_u8 = t.empty_like(kmers_u8)
            self.dist.broadcast(kmers_u8, src=0)
        packed = sh.search_kmers_hits(kmers_u8, q_offsets, n_queries, min_kmers, max_query_kmers)
        if self.dist is None:
            return packed[None]
        gathered = t.empty((self.world_size * packed.shape[0],), dtype=packed.dtype, device=sh.device)
        self.dist.all_gather_into_tensor(gathered, packed)  # exchange 2: per-shard hits
        return gathered.view(self.world_size, packed.shape[0])

    def to_global(self, gathered, n_queries, col_offsets, q=0):
        """Host-side merge of one query's gathered hits into global ascending colours."""
        cap = self.shard.cap
        n, cols, vals = unpack_hits(gathered.cpu().numpy(), n_queries, cap)
        if (n[:, q] > cap).any():
            raise OverflowError("hit capacity %d exceeded (%d hits on one shard)" % (cap, int(n[:, q].max())))
        return merge_shard_hits(n[:, q], cols[:, q], vals[:, q], col_offsets)
