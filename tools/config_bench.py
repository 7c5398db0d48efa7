"""GPU measurement (not the bench.py contract): the other BASELINE configs on ONE 50 000-column shard of
the m = 25 M index (what every GPU of the 8-way column-sharded ENA-scale index does), device-resident
inputs, CUDA events, distinct inputs per repetition:
  config 3  one 1 Mbp query (999 970 k-mers), exact (min_kmers = U)
  config 4  the same query at score >= 0.4, and a 10 000-k-mer query at 0.4
  config 5  1 000 queries x 1 000 k-mers in one launch: independent, and windows of one 100 kbp sequence
Prints one JSON line per case (lookups/s, algorithmic GB/s = k-mers * h * 6 250 B / time, fraction of the
measured copy peak).  Usage: python tools/config_bench.py [--m 25000000]"""
import argparse
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bigsi_b200 as B  # noqa: E402
from bigsi_b200.sharded import DeviceShard  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=25_000_000)
ap.add_argument("--cols", type=int, default=50_000)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
K, H = 31, 3
peak = 6533.5
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
ix = B.DeviceIndex(args.m, args.cols, col_offset=150_000)
ix.fill_synthetic(0, 1, [150_000, 150_001, 199_999, 175_000], [0xFFFFFFFF] * 3 + [int(0.95 * 2 ** 32)])
shard = DeviceShard(ix, K, H, cap=1024)
dev = shard.device
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
row_bytes = (args.cols + 7) // 8


def run(name, batches, qoff, mins, nq):
    """batches: list of uint8 [U, K] device tensors (distinct inputs); one launch each."""
    d_qoff = torch.tensor(qoff, dtype=torch.int64, device=dev)
    d_min = torch.tensor(mins, dtype=torch.int32, device=dev)
    maxq = int(np.diff(qoff).max())
    for b in batches[:2]:
        shard.search_kmers_hits(b, d_qoff, nq, d_min, maxq)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(args.reps):
        out = shard.search_kmers_hits(batches[r % len(batches)], d_qoff, nq, d_min, maxq)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    U = batches[0].shape[0]
    info = ix.info()
    gbs = U * H * row_bytes / (ms * 1e-3) / 1e9
    print(json.dumps({"case": name, "kmers_per_launch": U, "queries_per_launch": nq, "ms_per_launch": ms,
                      "lookups_per_s": U / (ms * 1e-3), "algorithmic_GBps": gbs, "frac_of_measured_copy_peak": gbs / peak,
                      "kernels_per_launch": 1 if info["last_fused"] & 1 else 2, "grid": info["last_grid"],
                      "n_slices": info["last_n_slices"], "stages": info["last_n_stages"],
                      "distinct_row_tuples_gathered_once": info["last_unique_kmers"],
                      "dram_bytes_moved_estimate_GB": (((H + 1) * info["last_unique_kmers"] + U) * row_bytes / 1e9
                                                       if info["last_unique_kmers"] else U * H * row_bytes / 1e9),
                      "first_query_hits": int(out[0].item())}))


def kmers_of_sequence(n_bases, seed):
    rng = np.random.default_rng(seed)
    s = acgt[rng.integers(0, 4, size=n_bases)]
    win = np.lib.stride_tricks.sliding_window_view(s, K)
    return torch.from_numpy(np.ascontiguousarray(win)).to(dev)  # random 31-mers: distinct with overwhelming probability


mega = [kmers_of_sequence(1_000_000, 2 + i) for i in range(3)]
U = mega[0].shape[0]
run("config3: 1 Mbp query, exact (min_kmers = U)", mega, [0, U], [U], 1)
run("config4: 1 Mbp query, score >= 0.4", mega, [0, U], [math.ceil(U * 0.4)], 1)
small = [kmers_of_sequence(10_030, 100 + i) for i in range(16)]
run("config4: 10 000-k-mer query, score >= 0.4", small, [0, 10_000], [4000], 1)
del mega
Q, L = 1000, 1000
qoff = list(range(0, Q * L + 1, L))
rng = np.random.default_rng(5)
indep = [torch.from_numpy(acgt[rng.integers(0, 4, size=(Q * L, K))]).to(dev) for _ in range(3)]
run("config5: 1 000 x 1 000 k-mers, independent queries", indep, qoff, [L] * Q, Q)
ix.set_option("batch_reuse", 0)
run("config5: 1 000 x 1 000 k-mers, independent queries, option batch_reuse = 0 (no de-duplication pass)", indep, qoff, [L] * Q, Q)
ix.set_option("batch_reuse", 1)
del indep
shared = []
for i in range(3):
    base = acgt[rng.integers(0, 4, size=100_000 + K)]
    win = np.lib.stride_tricks.sliding_window_view(base, K)
    starts = rng.integers(0, 100_000 - L, size=Q)
    shared.append(torch.from_numpy(np.ascontiguousarray(np.concatenate([win[s : s + L] for s in starts]))).to(dev))
run("config5: 1 000 x 1 000 k-mers, windows of one 100 kbp sequence (rows shared between queries)", shared, qoff,
    [math.ceil(L * 0.4)] * Q, Q)
ix.set_option("batch_reuse", 0)
run("config5: 1 000 x 1 000 k-mers, windows of one 100 kbp sequence, option batch_reuse = 0", shared, qoff,
    [math.ceil(L * 0.4)] * Q, Q)
ix.close()
