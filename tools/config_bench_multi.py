"""GPU measurement + parity gate (not the bench.py contract): BASELINE configs 3, 4 and 5 on the column-sharded index,
one process per GPU (torchrun), N = world x 50 000 columns of the m = 25 M index (world = 8: the ENA-scale N = 400 000
of configs[2]):
  config 3  one 1 Mbp query (999 970 k-mers), exact (min_kmers = U)
  config 4  the same query at score >= 0.4
  config 5  1 000 queries x 1 000 k-mers in one launch: independent, and windows of one 100 kbp sequence
These are long launches (3-4 ms per shard), so the two exchanges are plain collectives around the kernel
(ShardedSearcher.search_step: NCCL broadcast of the k-mer bytes, one launch per rank, NCCL all-gather of the packed
hit lists); the in-kernel exchange is for the latency-bound single short query (bench.py).
Parity on every rank before timing: its shard's counts of two 10 000-k-mer pieces of the megabase query against the
oracle (rows regenerated on the CPU from the synthetic index's pure function with GLOBAL column ids), additivity of
the pieces, and the all-gathered hit lists of every shard against the planted columns.
Prints one JSON line per case on rank 0: lookups/s (whole job), algorithmic GB/s per GPU, fraction of the measured copy
peak, max over ranks of the CUDA-event time.
Usage: torchrun --nproc-per-node N tools/config_bench_multi.py   (or plain python for N = 1)"""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import bigsi_b200 as B  # noqa: E402
from bigsi_b200.sharded import DeviceShard, ShardedSearcher, unpack_hits  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=25_000_000)
ap.add_argument("--cols", type=int, default=50_000)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--no-oracle", action="store_true")
args = ap.parse_args()
K, H, CAP = 31, 3, 1024
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
peak = 6533.5
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
cols = args.cols
pc, pt = bench.planted_columns(world, cols)
ix = B.DeviceIndex(args.m, cols, col_offset=rank * cols, device=local)
ix.fill_synthetic(0, 1, pc, pt)
shard = DeviceShard(ix, K, H, cap=CAP)
searcher = ShardedSearcher(shard, dist if world > 1 else None, world, rank)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
row_bytes = (cols + 7) // 8
ones = [0, 1, cols - 1]          # all-ones planted columns of every shard (local ids)
graded = cols // 2               # density 0.95: count ~ 0.857 U


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def kmers_of_sequence(n_bases, seed):
    rng = np.random.default_rng(seed)
    s = acgt[rng.integers(0, 4, size=n_bases)]
    return np.ascontiguousarray(np.lib.stride_tricks.sliding_window_view(s, K))  # random 31-mers: distinct w.o.p.


def run(name, batches, qoff, mins, nq, check):
    d_qoff = torch.tensor(qoff, dtype=torch.int64, device=dev)
    d_min = torch.tensor(mins, dtype=torch.int32, device=dev)
    maxq = int(np.diff(qoff).max())
    for b in batches[:2]:
        g = searcher.search_step(b, d_qoff, d_min, nq, maxq)
    torch.cuda.synchronize()
    n, hc, hv = unpack_hits(g.cpu().numpy(), nq, CAP)
    check(n, hc, hv)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(args.reps):
        searcher.search_step(batches[r % len(batches)], d_qoff, d_min, nq, maxq)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    U = batches[0].shape[0]
    gbs = U * H * row_bytes / (ms * 1e-3) / 1e9
    if rank == 0:
        info = ix.info()
        print(json.dumps({"case": name, "n_gpus": world, "total_columns": world * cols, "kmers_per_launch": U,
                          "queries_per_launch": nq, "ms_per_launch_max_over_ranks": ms,
                          "lookups_per_s_whole_job": world * U / (ms * 1e-3), "algorithmic_GBps_per_gpu": gbs,
                          "frac_of_measured_copy_peak": gbs / peak, "grid": info["last_grid"], "exchange":
                          "none" if world == 1 else "NCCL broadcast (k-mer bytes) + all-gather (packed hits) around one launch per rank",
                          "parity": "checked"}), flush=True)
    return ms


# ---- parity of this rank's shard on the megabase query -----------------------------------------------------------
mega_np = [kmers_of_sequence(1_000_000, 2 + i) for i in range(2)]
U = mega_np[0].shape[0]
if not args.no_oracle:
    from oracle import oracle as O

    oix = O.OracleIndex(K, args.m, H, cols, synth=O.SynthSpec(0, 1, pc, pt), col_offset=rank * cols)
    full = ix.search_kmers(mega_np[0], K, H)[0].astype(np.int64)
    acc = np.zeros_like(full)
    for i, p0 in enumerate(range(0, U, 10_000)):
        part = mega_np[0][p0: p0 + 10_000]
        c = ix.search_kmers(part, K, H)[0].astype(np.int64)
        if i in (0, 57):
            exp = oix.counts([bytes(r).decode() for r in part]).astype(np.int64)
            assert np.array_equal(c, exp), "rank %d: piece %d differs from the oracle" % (rank, i)
        acc += c
    assert np.array_equal(acc, full), "rank %d: additivity" % rank
    assert all(full[c] == U for c in ones) and abs(full[graded] - 0.95 ** 3 * U) < 0.01 * U
mega = [torch.from_numpy(a).to(dev) for a in mega_np]


def check_mega(thr_graded):
    def f(n, hc, hv):
        for r in range(world):  # every rank holds every shard's hit list (LOCAL colours of shard r)
            got = dict(zip(hc[r, 0, : int(n[r, 0])].tolist(), hv[r, 0, : int(n[r, 0])].tolist()))
            want = sorted(ones + ([graded] if thr_graded else []))
            assert sorted(got) == want, "rank %d sees shard %d hits %r" % (rank, r, sorted(got)[:8])
            assert all(got[c] == U for c in ones)
    return f


run("config3: 1 Mbp query, exact (min_kmers = U)", mega, [0, U], [U], 1, check_mega(False))
run("config4: 1 Mbp query, score >= 0.4", mega, [0, U], [math.ceil(U * 0.4)], 1, check_mega(True))
del mega
Q, L = 1000, 1000
qoff = list(range(0, Q * L + 1, L))
rng = np.random.default_rng(5)


def check_batch(thr_graded):
    def f(n, hc, hv):
        for r in range(world):
            for q in (0, 1, 499, 999):
                got = sorted(hc[r, q, : int(n[r, q])].tolist())
                assert got == sorted(ones + ([graded] if thr_graded else [])), (rank, r, q, got[:8])
    return f


indep = [torch.from_numpy(acgt[rng.integers(0, 4, size=(Q * L, K))]).to(dev) for _ in range(2)]
run("config5: 1 000 x 1 000 k-mers, independent queries, exact", indep, qoff, [L] * Q, Q, check_batch(False))
del indep
shared = []
for i in range(2):
    base = acgt[rng.integers(0, 4, size=100_000 + K)]
    win = np.lib.stride_tricks.sliding_window_view(base, K)
    starts = rng.integers(0, 100_000 - L, size=Q)
    shared.append(torch.from_numpy(np.ascontiguousarray(np.concatenate([win[s: s + L] for s in starts]))).to(dev))
run("config5: 1 000 x 1 000 k-mers, windows of one 100 kbp sequence (rows shared between queries), score >= 0.4", shared, qoff,
    [math.ceil(L * 0.4)] * Q, Q, check_batch(True))
barrier()
ix.close()
if world > 1:
    dist.destroy_process_group()
