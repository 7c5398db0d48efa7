"""Diagnostics: back-to-back single-query searches (the bench's value arm, nothing else) under several option sets.

usage: python tools/stream_probe.py [--m M] [--cols N] [--kmers U] [--queries Q] [--timeline] "k=v,k=v" "k=v" ...
An empty string "" is the default option set.  Prints one line per set: us/query, achieved GB/s."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import bigsi_b200  # noqa: E402
from bigsi_b200.sharded import DeviceShard, unpack_hits  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=25_000_000)
    ap.add_argument("--cols", type=int, default=50_000)
    ap.add_argument("--kmers", type=int, default=10_000)
    ap.add_argument("--queries", type=int, default=640)
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("sets", nargs="*", default=[""])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    K, H, U = bench.K, bench.H, a.kmers
    pc, pt = bench.planted_columns(1, a.cols)
    queries = bench.make_queries(64, U)
    d_queries = torch.from_numpy(queries).to(dev)
    d_qoff = torch.tensor([0, U], dtype=torch.int64, device=dev)
    d_min = torch.tensor([U], dtype=torch.int32, device=dev)
    row_bytes = (a.cols + 7) // 8
    for spec in a.sets:
        index = bigsi_b200.DeviceIndex(a.m, a.cols, device=0)
        index.fill_synthetic(0, 1, pc, pt)
        index.set_option("inputs_ready", 1)
        for kv in [x for x in spec.split(",") if x]:
            key, val = kv.split("=")
            index.set_option(key, int(val))
        shard = DeviceShard(index, K, H, cap=1024)

        def q(i):
            return shard.search_kmers_hits_stream(d_queries[i % 64], U)

        g = q(0)
        index.flush()
        torch.cuda.synchronize()
        n, hc, hv = unpack_hits(g.cpu().numpy(), 1, 1024)
        ok = sorted(hc[0, 0, : int(n[0, 0])].tolist()) == [0, 1, a.cols - 1]
        for i in range(200):
            q(i)
        torch.cuda.synchronize()
        best = None
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.queries):
                g = q(i)
            index.flush()
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / a.queries
            best = us if best is None else min(best, us)
        n, hc, hv = unpack_hits(g.cpu().numpy(), 1, 1024)
        ok = ok and sorted(hc[0, 0, : int(n[0, 0])].tolist()) == [0, 1, a.cols - 1]
        info = index.info()
        print("[%s] us/query %.2f  %.0f GB/s  hits_ok=%s  grid=%d smem=%d stages=%d reduce_grid=%d fused=%d"
              % (spec, best, U * H * row_bytes / best / 1e3, ok, info["last_grid"], info["last_smem_bytes"], info["last_n_stages"],
                 info["last_reduce_grid"], info["last_fused"]), flush=True)
        if a.timeline:
            bench.dump_timeline(index, lambda i: q(i), torch.cuda.synchronize, 0)
        index.close()
        del shard, index
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
