"""GPU experiment driver (not part of the product): steady-state time per query of the single-kernel
search path (raw k-mers -> hits) under different options.  Usage: python tools/sweep_solo.py [--configs JSON]"""
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
ap.add_argument("--reps", type=int, default=400)
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
kms = [torch.from_numpy(acgt[rng.integers(0, 4, size=(args.kmers, K))]).to(dev) for _ in range(NQ)]
d_min = torch.tensor([args.kmers], dtype=torch.int32, device=dev)
d_q = torch.tensor([0, args.kmers], dtype=torch.int64, device=dev)
algo = args.kmers * H * ((args.cols + 7) // 8)
configs = [{}, {"pool_pct": 0}, {"pool_pct": 20}, {"pool_pct": 30}, {"solo": 0}, {"n_stages": 8}, {"n_stages": 10},
           {"merge_chunk_bytes": 64}, {"merge_chunk_bytes": 32}, {"merge_chunk_bytes": 96}]
if args.configs:
    configs = json.loads(args.configs)
defaults = {"pool_pct": 12, "solo": 1, "n_stages": 0, "merge_chunk_bytes": 0, "tile_bytes": 0, "grid": 0, "prehash": 1, "fuse_merge": 1}
for cfg in configs:
    for k, v in defaults.items():
        ix.set_option(k, cfg.get(k, v))
    try:
        for i in range(20):
            shard.search_kmers_hits(kms[i % NQ], d_q, 1, d_min, args.kmers)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.reps):
            shard.search_kmers_hits(kms[i % NQ], d_q, 1, d_min, args.kmers)
        e1.record()
        torch.cuda.synchronize()
        step_us = 1e3 * e0.elapsed_time(e1) / args.reps
        ix.set_option("timing", 1)
        for i in range(100):
            shard.search_kmers_hits(kms[i % NQ], d_q, 1, d_min, args.kmers)
        torch.cuda.synchronize()
        f, mg, n = ix.timing_collect()
        ix.set_option("timing", 0)
        info = ix.info()
        print(json.dumps({"cfg": cfg, "step_us": step_us, "kernel_us": 1e3 * f / n, "GBps_step": algo / (step_us * 1e-6) / 1e9,
                          "GBps_kernel": algo / (f / n * 1e-3) / 1e9, "stages": info["last_n_stages"], "fused": info["last_fused"]}),
              flush=True)
    except Exception as e:  # keep sweeping
        print(json.dumps({"cfg": cfg, "error": str(e)}), flush=True)
        ix.set_option("timing", 0)
