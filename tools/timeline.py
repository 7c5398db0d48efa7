"""GPU experiment (not part of the product): per-CTA timeline of the fused kernel from the
globaltimer stamps recorded with option debug_flags bit 1.  Usage: python tools/timeline.py"""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bigsi_b200 as B  # noqa: E402
from bigsi_b200 import _lib  # noqa: E402
from bigsi_b200.sharded import DeviceShard  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=25_000_000)
ap.add_argument("--cols", type=int, default=50_000)
ap.add_argument("--kmers", type=int, default=10_000)
ap.add_argument("--flags", type=int, default=2)
ap.add_argument("--opt", action="append", default=[], help="index option key=value (repeatable)")
args = ap.parse_args()

K, H = 31, 3
ix = B.DeviceIndex(args.m, args.cols)
ix.fill_synthetic(0, 1, [0], [0xFFFFFFFF])
shard = DeviceShard(ix, K, H)
dev = shard.device
rng = np.random.default_rng(0)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
kms = [torch.from_numpy(acgt[rng.integers(0, 4, size=(args.kmers, K))]).to(dev) for _ in range(16)]
d_min = torch.tensor([args.kmers], dtype=torch.int32, device=dev)
d_q = torch.tensor([0, args.kmers], dtype=torch.int64, device=dev)
out = torch.empty((1, args.cols + 64), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ix.set_option("debug_flags", args.flags)
for kv in args.opt:
    key, val = kv.split("=")
    ix.set_option(key, int(val))
names = ["entry", "prod_first_issue", "first_slot_landed", "last_slot_consumed", "flushed", "prod_last_issue",
         "past_grid_barrier", "merge_done", "past_pdl_wait", "hash_done", "merge_loaded", "merge_in_smem",
         "merge_expanded", "merge_stage_issue"]
NS = 16
for rep in range(6):
    shard.search_kmers_hits(kms[rep], d_q, 1, d_min, args.kmers)
    torch.cuda.synchronize()
    grid = ix.info()["last_grid"]
    buf = np.zeros(grid * NS, dtype=np.uint64)
    _lib.check(_lib.lib().bigsi_b200_index_debug_read(ix.handle, ctypes.c_void_p(buf.ctypes.data), buf.size))
    ts = buf.reshape(grid, NS).astype(np.int64)
    t0 = ts[:, 0].min()
    rel = (ts[:, :NS] - t0) / 1e3
    if rep < 2:
        continue
    print("launch %d: grid=%d  (microseconds since the first CTA entered the kernel; min / median / max over CTAs)" % (rep, grid))
    for i, n in enumerate(names):
        col = rel[:-1, i]  # the last CTA may hold a short remainder slice
        col = col[ts[:-1, i] > 0]
        if col.size == 0:
            continue
        print("  %-20s %7.2f %7.2f %7.2f" % (n, col.min(), np.median(col), col.max()))
    mhz = (ts[:, 15] - ts[:, 14]) / np.maximum(ts[:, 7] - ts[:, 0], 1) * 1e3
    print("  SM clock over the kernel (clock64 / globaltimer): min %.0f median %.0f max %.0f MHz" % (mhz.min(), np.median(mhz), mhz.max()))
    print("  kernel span (first entry -> last flush): %.2f us; -> merge done: %.2f us" % (rel[:, 4].max(), rel[:, 7].max()))
