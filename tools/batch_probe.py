"""GPU experiment: one config-5 batch (shared windows / independent) for an ncu launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bigsi_b200 as B
from bigsi_b200.sharded import DeviceShard
K, H, Q, L = 31, 3, 1000, 1000
ix = B.DeviceIndex(25_000_000, 50_000)
ix.fill_synthetic(0, 1, [0], [0xFFFFFFFF])
shard = DeviceShard(ix, K, H)
dev = shard.device
rng = np.random.default_rng(5)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
base = acgt[rng.integers(0, 4, size=100_000 + K)]
win = np.lib.stride_tricks.sliding_window_view(base, K)
starts = rng.integers(0, 100_000 - L, size=Q)
shared = torch.from_numpy(np.ascontiguousarray(np.concatenate([win[s: s + L] for s in starts]))).to(dev)
indep = torch.from_numpy(acgt[rng.integers(0, 4, size=(Q * L, K))]).to(dev)
d_min = torch.full((Q,), 400, dtype=torch.int32, device=dev)
d_q = torch.arange(0, Q * L + 1, L, dtype=torch.int64, device=dev)
for rep in range(2):
    shard.search_kmers_hits(shared, d_q, Q, d_min, L)
    shard.search_kmers_hits(indep, d_q, Q, d_min, L)
torch.cuda.synchronize()
print("done", ix.info()["last_unique_kmers"])
