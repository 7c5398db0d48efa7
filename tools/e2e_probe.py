"""GPU experiment (not part of the product): where the end-to-end time of one host-buffer search goes.
Times the Python wrapper, the bare ctypes call and the staged (copy) path on the config-2 index."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bigsi_b200 as B  # noqa: E402
from bigsi_b200 import _lib  # noqa: E402

K, H, U, CAP = 31, 3, 10_000, 1024
ix = B.DeviceIndex(25_000_000, 50_000)
ix.fill_synthetic(0, 1, [0, 1, 49_999], [0xFFFFFFFF] * 3)
rng = np.random.default_rng(0)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
q = torch.from_numpy(acgt[rng.integers(0, 4, size=(64, U, K))]).pin_memory()
qs = [q[i].numpy() for i in range(64)]
qp = [np.array(acgt[rng.integers(0, 4, size=(U, K))]) for i in range(64)]  # pageable
mk = np.array([U], dtype=np.uint32)
L = _lib.lib()


def timeit(fn, n=400):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / n


print("wrapper, pinned, zero-copy      %.1f us" % timeit(lambda i: ix.search_kmers_hits(qs[i % 64], K, H, mk, cap=CAP)))
print("wrapper, pageable, zero-copy    %.1f us" % timeit(lambda i: ix.search_kmers_hits(qp[i % 64], K, H, mk, cap=CAP)))
cols = np.empty(CAP, dtype=np.int32)
cnts = np.empty(CAP, dtype=np.uint32)
n = np.zeros(1, dtype=np.uint64)
qo = np.array([0, U], dtype=np.int64)
args = [(ix.handle, a.ctypes.data, qo.ctypes.data, 1, K, H, mk.ctypes.data, cols.ctypes.data, cnts.ctypes.data, CAP, n.ctypes.data)
        for a in qs]
f = L.bigsi_b200_search_kmers_hits
print("bare ctypes, pinned, zero-copy  %.1f us" % timeit(lambda i: f(*args[i % 64])))
assert n[0] == 3, n
ix.set_option("zero_copy", 0)
print("bare ctypes, pinned, staged     %.1f us" % timeit(lambda i: f(*args[i % 64])))
ix.set_option("zero_copy", 1)
ix.set_option("timing", 1)
for i in range(200):
    f(*args[i % 64])
fm, mm, nn = ix.timing_collect()
print("kernel (events) in the zero-copy path %.1f us" % (1e3 * fm / nn))
seq = bytes(acgt[rng.integers(0, 4, size=U + K - 1)])
seqs = [bytes(acgt[rng.integers(0, 4, size=U + K - 1)]) for _ in range(16)]
print("search_sequence (wrapper, 10 030-base sequence, threshold 1.0)  %.1f us" % timeit(lambda i: ix.search_sequence(seqs[i % 16], K, H, 1.0, cap=CAP)))
nh = np.zeros(1, dtype=np.uint64)
nu = np.zeros(1, dtype=np.uint64)
g = L.bigsi_b200_search_sequence
sargs = [(ix.handle, s, len(s), K, H, ctypes.c_double(1.0), cols.ctypes.data, cnts.ctypes.data, CAP, nh.ctypes.data, nu.ctypes.data)
         for s in seqs]
print("search_sequence (bare ctypes)  %.1f us" % timeit(lambda i: g(*sargs[i % 16])))
assert nh[0] == 3 and nu[0] == U, (nh, nu)
for i in range(200):
    g(*sargs[i % 16])
fm, mm, nn = ix.timing_collect()
print("query kernel (events) in the sequence path %.1f us" % (1e3 * fm / nn))
ix.set_option("timing", 0)
ix.set_option("zero_copy", 0)
print("search_sequence (bare ctypes, host round trip for U)  %.1f us" % timeit(lambda i: g(*sargs[i % 16])))
ix.set_option("zero_copy", 1)
# host-side cost of the call alone: an empty-ish query (k-mer count 1) through the same entry point
tiny = [bytes(acgt[rng.integers(0, 4, size=K)]) for _ in range(16)]
targs = [(ix.handle, s, len(s), K, H, ctypes.c_double(1.0), cols.ctypes.data, cnts.ctypes.data, CAP, nh.ctypes.data, nu.ctypes.data)
         for s in tiny]
print("search_sequence (bare ctypes, ONE k-mer: launch + front-end + kernel fixed costs)  %.1f us" % timeit(lambda i: g(*targs[i % 16])))
