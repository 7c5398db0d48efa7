"""CPU, world_size = 2, gloo: the column-sharded exchange logic of bigsi_b200.sharded
(broadcast of the query from rank 0, local search, one all-gather of packed hits, host merge into
global colours).  The local compute is injected: here an oracle-backed stand-in with the same
interface as DeviceShard (the product's DeviceShard is CUDA-only and is covered by the -m gpu
tests); what is under test is the sharding / exchange / merge code."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, H, M, N, CAP = 15, 3, 20_011, 1003, 512


class OracleShard:
    """Test stand-in for bigsi_b200.sharded.DeviceShard over oracle.OracleIndex (CPU tensors)."""

    def __init__(self, oix, col_offset, cap):
        self.torch = torch
        self.oix = oix
        self.k, self.h, self.m = oix.k, oix.h, oix.m
        self.num_cols = oix.num_cols
        self.col_offset = col_offset
        self.device = torch.device("cpu")
        self.cap = cap

    def search_kmers_hits(self, kmers_u8, q_offsets, n_queries, min_kmers, max_query_kmers=0):
        arr = kmers_u8.numpy()
        qo = q_offsets.numpy()
        buf = np.zeros(n_queries * (2 + 2 * self.cap), dtype=np.int32)
        n64 = buf[: 2 * n_queries].view(np.int64)
        cols = buf[2 * n_queries : 2 * n_queries + n_queries * self.cap].reshape(n_queries, self.cap)
        vals = buf[2 * n_queries + n_queries * self.cap :].reshape(n_queries, self.cap)
        for q in range(n_queries):
            km = [bytes(r).decode() for r in arr[qo[q] : qo[q + 1]]]
            cnt = self.oix.counts(km) if km else np.zeros(self.num_cols, dtype=np.int32)
            hit = np.nonzero(cnt >= int(min_kmers[q]))[0][::-1]  # unspecified order: hand them over reversed
            n64[q] = len(hit)
            cols[q, : min(len(hit), self.cap)] = hit[: self.cap]
            vals[q, : min(len(hit), self.cap)] = cnt[hit[: self.cap]]
        return torch.from_numpy(buf)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bigsi_b200.sharded import ShardedSearcher, shard_columns
        from oracle import oracle as O

        rng = np.random.default_rng(5)  # same matrix and queries on both ranks
        rows = np.packbits(rng.random((M, N)) < 0.8, axis=1)
        full = O.OracleIndex(K, M, H, N, rows=rows)
        ranges = shard_columns(N, world)
        a, b = ranges[rank]
        bits = np.unpackbits(rows, axis=1)[:, a:b]
        shard = OracleShard(O.OracleIndex(K, M, H, b - a, rows=np.packbits(bits, axis=1)), a, CAP)
        searcher = ShardedSearcher(shard, dist, world, rank)
        acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
        lens = [40, 0, 7]
        qoff = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
        kmers = acgt[rng.integers(0, 4, size=(sum(lens), K))]
        mins = torch.tensor([24, 1, 5], dtype=torch.int32)
        # only rank 0 holds the real query; the other rank passes a same-shaped placeholder
        mine = torch.from_numpy(kmers.copy() if rank == 0 else np.zeros_like(kmers))
        gathered = searcher.search_step(mine, qoff, mins, len(lens), max(lens))
        assert tuple(gathered.shape) == (world, len(lens) * (2 + 2 * CAP))
        ok = True
        for q in range(len(lens)):
            km = [bytes(r).decode() for r in kmers[qoff[q] : qoff[q + 1]]]
            cnt = full.counts(km) if km else np.zeros(N, dtype=np.int32)
            exp = np.nonzero(cnt >= int(mins[q]))[0]
            cols, vals = searcher.to_global(gathered, len(lens), [r[0] for r in ranges], q)
            ok &= np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_world2_gloo():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(ret) == {0: True, 1: True}
