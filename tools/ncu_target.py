"""GPU experiment (not part of the product): a short workload for `ncu` captures -- the single-kernel
query path on the m = 2.5 M twin of BASELINE config 2 (same row length and random gather; the 157 GB
index cannot be replayed), the query front-end, the build path's transpose and the score=True kernels.
Usage (under gpurun): ncu --set full ... python tools/ncu_target.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bigsi_b200 as B  # noqa: E402

K, H, U, CAP = 31, 3, 10_000, 1024
m, cols = 2_500_000, 50_000
ix = B.DeviceIndex(m, cols)
ix.fill_synthetic(0, 1, [0, 1, cols - 1], [0xFFFFFFFF] * 3)
rng = np.random.default_rng(0)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
mk = np.array([U], dtype=np.uint32)
for i in range(6):  # fused_query<COUNTS,3,solo>: k-mers in, hits out
    q = acgt[rng.integers(0, 4, size=(U, K))]
    c, v, n = ix.search_kmers_hits(q, K, H, mk, cap=CAP)[0]
    assert sorted(c.tolist()) == [0, 1, cols - 1]
for i in range(4):  # dedup_windows_kernel + fused_query reading U from the device
    seq = acgt[rng.integers(0, 4, size=U + K - 1)].tobytes()
    c, v, n, u = ix.search_sequence(seq, K, H, 1.0, cap=CAP)
    assert u == U and sorted(c.tolist()) == [0, 1, cols - 1]
pres = ix.sequence_presence(seq, K, H, [0, 1, 7, cols - 1])  # hash_windows_kernel + presence_kernel
assert pres.shape == (4, U)
ix.close()
# transpose_blooms_kernel: 4 096 Bloom filters of 2.5 M bits -> columns
nb = 4096
ix = B.DeviceIndex(m, 0, col_capacity=nb)
blooms = rng.integers(0, 256, size=(nb, (m + 7) // 8), dtype=np.uint8)
ix.build_columns(0, blooms, m)
# bloom_set_bits_kernel
B.index.bloom_kmers(acgt[rng.integers(0, 4, size=(200_000, K))], K, H, 25_000_000)
ix.close()
print("ncu target done")
