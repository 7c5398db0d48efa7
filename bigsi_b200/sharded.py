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
        # single-query searches are "streamed" (include/bigsi_b200.h): stage 2 of a query runs while the next query's rows
        # stream, so the output buffers of consecutive calls must differ -- a ring of 8
        self._hit_ring = {}
        self._hit_turn = 0

    def _hit_buffer(self, n_queries):
        t = self.torch
        ring = self._hit_ring.get(n_queries)
        if ring is None:
            if len(self._hit_ring) > 4:
                self._hit_ring.clear()
            ring = [t.empty((n_queries * (2 + 2 * self.cap),), dtype=t.int32, device=self.device) for _ in range(8)]
            self._hit_ring[n_queries] = ring
        self._hit_turn = (self._hit_turn + 1) % 8
        return ring[self._hit_turn]

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
        """Fused gather-AND-count + threshold (one C-ABI call): same packed layout as hits(); the full
        count vectors are never written.  The returned buffer is one of a ring of 8."""
        buf = self._hit_buffer(n_queries)
        base = buf.data_ptr()
        self.index.query_hits_dev(rows.data_ptr(), q_offsets.data_ptr(), n_queries, rows.shape[0], self.h,
                                  min_kmers.data_ptr(), base + 8 * n_queries,
                                  base + 8 * n_queries + 4 * n_queries * self.cap, self.cap, base, self._stream(),
                                  max_query_kmers)
        return buf


    def search_kmers_hits(self, kmers_u8, q_offsets, n_queries, min_kmers, max_query_kmers=0):
        """The whole path from raw k-mers uint8 [U, k] on the device: canonical + murmur3 in the kernel
        prologue -> gather-AND-count -> merge + threshold (one query: gather kernel + overlapped reduce
        kernel; a batch: one kernel with a grid barrier).  Same packed layout as hits(); the returned
        buffer is one of a ring of 8 and stays valid for the next 7 calls."""
        buf = self._hit_buffer(n_queries)
        base = buf.data_ptr()
        self.index.query_kmers_hits_dev(kmers_u8.data_ptr(), self.k, q_offsets.data_ptr(), n_queries, kmers_u8.shape[0],
                                        self.h, min_kmers.data_ptr(), base + 8 * n_queries,
                                        base + 8 * n_queries + 4 * n_queries * self.cap, self.cap, base, self._stream(),
                                        max_query_kmers)
        return buf


    def search_kmers_hits_stream(self, kmers_u8, min_kmers):
        """ONE query from raw k-mers uint8 [U, k] on the device, DEFERRED: same packed layout as hits() for one query,
        but the returned buffer (one of a ring of 8) is complete in stream order only after the NEXT streamed
        search on this shard or after flush() -- stage 2 of a query rides in the gather kernel of its successor."""
        buf = self._hit_buffer(1)
        base = buf.data_ptr()
        self.index.query_kmers_hits_stream_dev(kmers_u8.data_ptr(), self.k, kmers_u8.shape[0], self.h, int(min_kmers), base + 8,
                                               base + 8 + 4 * self.cap, self.cap, base, self._stream())
        return buf

    def flush(self):
        self.index.flush()


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class FusedExchange:
    """The two exchanges of a column-sharded single-query search done by the query kernels themselves
    (include/bigsi_b200.h "column-sharded search ... WITHOUT per-query collectives"): rank 0's gather kernel
    pushes the k-mer bytes into the peers' inboxes over NVLink, every rank's stage 2 (merge team / flush kernel) publishes its hits
    into every rank's result blocks and waits for the others while the next query's gather kernel already
    runs.  No NCCL call per query; torch.distributed is only used once, to exchange the CUDA IPC handles."""

    def __init__(self, shard, world_size, rank, max_kmers, dist=None, peers=None):
        import ctypes

        from . import _lib

        self._ct, self._lib = ctypes, _lib
        self.shard, self.world, self.rank = shard, world_size, rank
        self.spec = shard.cap
        self._views = {}
        L = _lib.lib()
        handle = (ctypes.c_uint8 * 64)()
        _lib.check(L.bigsi_b200_exchange_create(shard.index.handle, world_size, rank, max_kmers * shard.k, self.spec, handle))
        # all scratch now: a search must never have to grow (and free) a buffer while other shards' kernels wait for it
        _lib.check(L.bigsi_b200_exchange_reserve(shard.index.handle, max_kmers, shard.k, shard.h))
        if world_size > 1 and peers is None:
            gathered = [None] * world_size
            dist.all_gather_object(gathered, bytes(handle))
            blob = b"".join(gathered)
            _lib.check(L.bigsi_b200_exchange_open(shard.index.handle, (ctypes.c_uint8 * len(blob)).from_buffer_copy(blob)))
            dist.barrier()

    @staticmethod
    def connect_local(exchanges):
        """Same-process wiring of `world` FusedExchange objects (one per handle; the handles may share a device)."""
        import ctypes

        from . import _lib

        arr = (ctypes.c_void_p * len(exchanges))(*[e.shard.index.handle.value for e in exchanges])
        for e in exchanges:
            _lib.check(_lib.lib().bigsi_b200_exchange_open_local(e.shard.index.handle, arr))

    def _view(self, ptr, stride):
        if not ptr:
            return None
        view = self._views.get(ptr)
        if view is None:  # eight generations of result blocks rotate: build each torch view once
            t = self.shard.torch
            blocks = t.as_tensor(_DevArray(ptr, (self.world, stride // 4), "<i4"), device=self.shard.device)
            view = blocks[:, 2 : 4 + 2 * self.spec]  # drop the sequence word: [n (2 x int32) | cols | counts]
            self._views[ptr] = view
        return view

    def search(self, kmers_u8, n_kmers, min_kmers, stream=None):
        """One query (rank 0's k-mers decide; a device tensor, or any 16-byte aligned device-addressable
        address as an int).  Returns int32 [world, 2 + 2*spec] (LOCAL colours), the packed layout of
        DeviceShard.hits for one query: a view of library memory, DEFERRED -- complete in stream order after the
        next search() or after flush() (on every rank) -- and valid until four more searches have been issued."""
        ct = self._ct
        ptr, stride = ct.c_void_p(0), ct.c_uint64(0)
        d_k = 0
        if self.rank == 0 and kmers_u8 is not None:
            d_k = kmers_u8 if isinstance(kmers_u8, int) else kmers_u8.data_ptr()
        self._lib.check(self._lib.lib().bigsi_b200_exchange_search_dev(
            self.shard.index.handle, d_k, n_kmers, self.shard.k, self.shard.h, int(min_kmers),
            self.shard._stream() if stream is None else stream, ct.byref(ptr), ct.byref(stride)))
        return self._view(ptr.value, stride.value)

    def flush(self):
        """Stage 2 of the last search as a kernel of its own (SPMD: every rank calls it)."""
        self.shard.index.flush()

    def enable_host_results(self, on=True):
        """The all-gathered hit lists of the searches launched while this is on are also written into mapped host
        memory by the kernels themselves (no copy in the stream); see wait_host.  Per rank; costs the query's last
        stage-2 CTA a system fence over PCIe, so ranks that do not read results on the host leave it off."""
        self._lib.check(self._lib.lib().bigsi_b200_exchange_host_results(self.shard.index.handle, 1 if on else 0))

    def last_seq(self):
        s = self._ct.c_uint64(0)
        self._lib.check(self._lib.lib().bigsi_b200_exchange_last_seq(self.shard.index.handle, self._ct.byref(s)))
        return s.value

    def wait_host(self, seq):
        """Blocks until search number `seq` (last_seq() right after its search()) is complete on this rank and returns
        its result as a numpy int32 view [world, 2 + 2*spec] of HOST memory (the packed layout of search()), valid
        until 8 more searches.  Flushes the search when it is the newest one (then every rank must flush too)."""
        ct = self._ct
        ptr, stride = ct.c_void_p(0), ct.c_uint64(0)
        self._lib.check(self._lib.lib().bigsi_b200_exchange_wait_host(self.shard.index.handle, seq, ct.byref(ptr), ct.byref(stride)))
        n = self.world * stride.value // 4
        arr = np.ctypeslib.as_array((ct.c_int32 * n).from_address(ptr.value)).reshape(self.world, stride.value // 4)
        return arr[:, 2: 4 + 2 * self.spec]

    def wait_ns(self):
        """(ns this rank's stage-2 code waited for the other shards since the last call, queries launched)."""
        ct = self._ct
        w, q = ct.c_uint64(0), ct.c_uint64(0)
        self._lib.check(self._lib.lib().bigsi_b200_exchange_wait_ns(self.shard.index.handle, ct.byref(w), ct.byref(q)))
        return w.value, q.value

    def close(self):
        self._lib.lib().bigsi_b200_exchange_destroy(self.shard.index.handle)


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

    def __init__(self, shard, dist=None, world_size=1, rank=0, fused_max_kmers=0):
        """fused_max_kmers > 0 (CUDA shards only): single-query searches of up to that many k-mers use
        the in-kernel exchange (FusedExchange) instead of a broadcast and an all-gather per query."""
        self.shard = shard
        self.dist = dist if world_size > 1 else None
        self.world_size = world_size
        self.rank = rank
        self.torch = shard.torch
        self.fused = None
        if fused_max_kmers and world_size > 1 and hasattr(shard, "index"):
            self.fused = FusedExchange(shard, world_size, rank, fused_max_kmers, dist=dist)
            self.fused_max_kmers = fused_max_kmers

    def search_one_fused(self, kmers_u8, n_kmers, min_kmers):
        """Single query through the in-kernel exchange; min_kmers is a host integer.  Same result layout
        as search_step for one query (FusedExchange.search): DEFERRED, see flush()."""
        return self.fused.search(kmers_u8, n_kmers, min_kmers)

    def flush(self):
        """Completes the last deferred search of this rank in stream order (every rank calls it)."""
        self.shard.index.flush()

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
