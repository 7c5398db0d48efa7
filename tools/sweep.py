"""GPU experiment driver (not part of the product): times the fused kernel under different launch
geometries on the BASELINE config-2 index.  Usage: python tools/sweep.py [--m M] [--cols N]"""
import argparse
import json
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
ap.add_argument("--kmers", type=int, default=10_000)
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--configs", default="")
args = ap.parse_args()

K, H = 31, 3
ix = B.DeviceIndex(args.m, args.cols)
ix.fill_synthetic(0, 1, [0], [0xFFFFFFFF])
shard = DeviceShard(ix, K, H)
dev = shard.device
rng = np.random.default_rng(0)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
NQ = 64
rows = [shard.hash(torch.from_numpy(acgt[rng.integers(0, 4, size=(args.kmers, K))]).to(dev)) for _ in range(NQ)]
rows_sorted = [torch.sort(r.reshape(-1))[0].reshape(-1, H) for r in rows]  # NOT the same k-mers: bandwidth probe only
rows_seq = [(torch.arange(args.kmers * H, dtype=torch.int32, device=dev) + (i * 300_007) % max(1, args.m - args.kmers * H)).reshape(-1, H)
            for i in range(NQ)]  # contiguous rows: a streaming read through the same kernel
d_q = torch.tensor([0, args.kmers], dtype=torch.int64, device=dev)
out = torch.empty((1, args.cols + 64), dtype=torch.int32, device=dev)
algo = args.kmers * H * ((args.cols + 7) // 8)

configs = [
    {},
    {"debug_flags": 1},
    {"n_stages": 2}, {"n_stages": 4}, {"n_stages": 8},
    {"tile_bytes": 3200}, {"tile_bytes": 2176}, {"tile_bytes": 1664}, {"tile_bytes": 1024}, {"tile_bytes": 512},
    {"grid": 296, "tile_bytes": 3200}, {"grid": 74}, {"grid": 296},
    {"sorted": 1}, {"sorted": 1, "debug_flags": 1},
]
if args.configs:
    configs = json.loads(args.configs)
keys = ("tile_bytes", "grid", "kmers_per_stage", "n_stages", "ctas_per_sm", "debug_flags")
for cfg in configs:
    for k in keys:
        ix.set_option(k, cfg.get(k, 0))
    rr = rows_sorted if cfg.get("sorted") else rows_seq if cfg.get("seq") else rows
    mode = cfg.get("mode", 0)
    try:
        for i in range(10):
            ix.query_dev(mode, rr[i % NQ].data_ptr(), d_q.data_ptr(), 1, args.kmers, H, out.data_ptr(), out.shape[1] * (1 if mode == 0 else 4),
                         torch.cuda.current_stream().cuda_stream, args.kmers)
        torch.cuda.synchronize()
        ix.set_option("timing", 1)
        for i in range(args.reps):
            ix.query_dev(mode, rr[i % NQ].data_ptr(), d_q.data_ptr(), 1, args.kmers, H, out.data_ptr(), out.shape[1] * (1 if mode == 0 else 4),
                         torch.cuda.current_stream().cuda_stream, args.kmers)
        torch.cuda.synchronize()
        f, mg, n = ix.timing_collect()
        ix.set_option("timing", 0)
        info = ix.info()
        print(json.dumps({"cfg": cfg, "fused_us": 1e3 * f / n, "merge_us": 1e3 * mg / n, "GBps": algo / (f / n * 1e-3) / 1e9,
                          "grid": info["last_grid"], "block": info["last_block"], "tile": info["last_tile_bytes"],
                          "stages": info["last_n_stages"], "G": info["last_kmers_per_stage"]}), flush=True)
    except Exception as e:  # keep sweeping
        print(json.dumps({"cfg": cfg, "error": str(e)}), flush=True)
        ix.set_option("timing", 0)
