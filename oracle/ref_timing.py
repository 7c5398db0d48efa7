"""BASELINE INFRASTRUCTURE ONLY -- times the UNMODIFIED reference package's own BIGSI.search (bigsi/graph/bigsi.py:
174-230: seq_to_kmers -> lookup (hash, storage.batch_get, per-k-mer AND) -> exact / inexact filter) on the bench's
synthetic index, single thread, as BASELINE.md section 3 line 1 plans it: the package is imported through the
stand-ins of oracle/ref_shims (mmh3 / bitarray / redis are not installable here), its rows live in the dict-backed
BaseStorage of oracle/ref_harness.py, and only the rows the query touches are stored (regenerated from the synthetic
index's pure function, as the device matrix is).  Used by bench.py's cpu_baseline leg when the package is present
(baseline/_ref/ on the GPU box); never by bigsi_b200/."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from oracle import ref_harness  # noqa: E402


def time_reference_search(m, cols, k, h, n_kmers, planted_cols, planted_thr, thresholds=(1.0, 0.4), seed=4242):
    """Returns {threshold: (seconds, lookups_per_s, hits)} for one sequence of n_kmers windows (all distinct)."""
    ref = ref_harness.load_reference()
    from bigsi.graph.metadata import SampleMetadata
    from bigsi.storage import get_storage

    cfg = ref_harness.dict_config("ref-timing", k, m, h)
    storage = get_storage(cfg)
    storage.delete_all()
    SampleMetadata(storage).add_samples(["s%d" % c for c in range(cols)])
    storage.set_integer("ksi:bloomfilter_size", m)
    storage.set_integer("ksi:num_hashes", h)
    storage.set_integer("number_of_rows", m)
    storage.set_integer("number_of_cols", cols)
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq_arr = acgt[rng.integers(0, 4, size=n_kmers + k - 1)]
    win = np.ascontiguousarray(np.lib.stride_tricks.sliding_window_view(seq_arr, k))
    assert len(np.unique(win, axis=0)) == n_kmers
    rows = np.unique(O.hash_kmers(win, k, h, m).reshape(-1))
    spec = O.SynthSpec(0, 1, list(planted_cols), list(planted_thr))
    data = spec.rows(rows, 0, cols)
    for r, b in zip(rows.tolist(), data):
        storage["%d:bitarray" % r] = b.tobytes()
    bigsi = ref.BIGSI(cfg)
    seq = seq_arr.tobytes().decode("ascii")
    out = {}
    for thr in thresholds:
        t0 = time.perf_counter()
        res = bigsi.search(seq, thr)
        dt = time.perf_counter() - t0
        out[thr] = (dt, n_kmers / dt, sorted(r["sample_name"] for r in res))
    storage.delete_all()
    return out


if __name__ == "__main__":
    import bench

    pc, pt = bench.planted_columns(1, 50_000)
    for thr, (dt, rate, hits) in time_reference_search(25_000_000, 50_000, 31, 3, int(sys.argv[1]) if len(sys.argv) > 1 else 500, pc, pt).items():
        print("threshold %.1f: %.3f s, %.0f lookups/s, hits %s" % (thr, dt, rate, hits))
