"""GPU experiment: per-CTA timeline of the generic (batch) kernel on BASELINE config 5 (1 000 x 1 000 k-mers)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bigsi_b200 as B
from bigsi_b200 import _lib
from bigsi_b200.sharded import DeviceShard
K, H, Q, L = 31, 3, int(os.environ.get("Q", 1000)), int(os.environ.get("L", 1000))
ix = B.DeviceIndex(25_000_000, 50_000)
ix.fill_synthetic(0, 1, [0], [0xFFFFFFFF])
shard = DeviceShard(ix, K, H)
dev = shard.device
rng = np.random.default_rng(0)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
kms = [torch.from_numpy(acgt[rng.integers(0, 4, size=(Q * L, K))]).to(dev) for _ in range(3)]
d_min = torch.full((Q,), L, dtype=torch.int32, device=dev)
d_q = torch.arange(0, Q * L + 1, L, dtype=torch.int64, device=dev)
for kv in sys.argv[1:]:
    key, val = kv.split("=")
    ix.set_option(key, int(val))
for rep in range(3):
    shard.search_kmers_hits(kms[rep], d_q, Q, d_min, L)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for rep in range(5):
    shard.search_kmers_hits(kms[rep % 3], d_q, Q, d_min, L)
e1.record(); torch.cuda.synchronize()
print("ms per launch %.3f" % (e0.elapsed_time(e1) / 5), {k: v for k, v in ix.info().items() if k.startswith("last_")})
ix.set_option("debug_flags", 2)
shard.search_kmers_hits(kms[0], d_q, Q, d_min, L)
torch.cuda.synchronize()
grid = ix.info()["last_grid"]
buf = np.zeros(grid * 16, dtype=np.uint64)
_lib.check(_lib.lib().bigsi_b200_index_debug_read(ix.handle, ctypes.c_void_p(buf.ctypes.data), buf.size))
ts = buf.reshape(grid, 16).astype(np.int64)
t0 = ts[:, 0].min()
names = {0: "entry", 8: "past_pdl_wait", 9: "hash_done", 1: "prod_first_issue", 2: "first_slot_landed", 5: "prod_last_issue", 3: "last_slot_consumed",
         4: "flushed(last segment)", 6: "past_grid_barrier", 7: "merge_done"}
names.update({13: "merge: last stage issue", 10: "merge: first batch loaded (last item)", 11: "merge: counters in smem (last item)", 12: "merge: expanded (last item)"})
d = ts[:, 12] - ts[:, 13]
d = d[(ts[:, 12] > 0) & (ts[:, 13] > 0)] / 1e3
if d.size: print("  last merge item per CTA: stage issue -> expanded: min %.1f median %.1f max %.1f us;  staging %.1f  counting %.1f  expansion %.1f (medians)" % (
    d.min(), np.median(d), d.max(), np.median((ts[:, 10] - ts[:, 13])[ts[:, 12] > 0]) / 1e3, np.median((ts[:, 11] - ts[:, 10])[ts[:, 12] > 0]) / 1e3,
    np.median((ts[:, 12] - ts[:, 11])[ts[:, 12] > 0]) / 1e3))
for i, n in names.items():
    col = ts[:, i]; col = (col[col > 0] - t0) / 1e3
    if col.size: print("  %-24s %9.1f %9.1f %9.1f" % (n, col.min(), np.median(col), col.max()))
