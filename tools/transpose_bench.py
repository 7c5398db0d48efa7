"""GPU measurement (not part of the product): transpose_blooms kernel on GPU-resident filters, CUDA events.
4 096 Bloom filters x 2.5 M bits -> 4 096 columns (1.28 GB read + 1.28 GB written algorithmically)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bigsi_b200 as B  # noqa: E402

m, n = 2_500_000, 4096
stride = (m + 255) // 256 * 32
d = torch.randint(0, 256, (n, stride), dtype=torch.uint8, device="cuda")
ix = B.DeviceIndex(m, 0, col_capacity=n)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ix.build_columns_dev(0, n, d.data_ptr(), stride, m, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 10
for _ in range(reps):
    ix.build_columns_dev(0, n, d.data_ptr(), stride, m, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
algo = 2 * n * (m / 8)
print("transpose_blooms: %d filters x %d bits: %.3f ms per launch, %.0f GB/s algorithmic (read + write)" % (n, m, ms, algo / ms / 1e6))
