"""Stress of the in-kernel sequence front-end on tiny indexes (diagnostics): repeated searches of a few short
sequences, every result compared with the oracle; prints the anomalies."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bigsi_b200 as B  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    bad = 0
    for (k, m, h, N) in ((3, 1000, 3, 2), (11, 20011, 3, 7), (31, 5003, 3, 37)):
        rows = rng.random((m, N)) < 0.8
        packed = np.packbits(rows, axis=1)
        ix = B.DeviceIndex(m, N)
        ix.upload_rows(0, packed)
        oix = O.OracleIndex(k, m, h, N, rows=packed)
        seqs = ["".join(rng.choice(list("ACGT"), size=n)) for n in (9, 40, 120, 300)]
        want = []
        for s in seqs:
            uk = O.unique_kmers(s, k)
            want.append((len(uk), oix.counts(uk) if uk else np.zeros(N, dtype=np.int64)))
        for it in range(400):
            j = it % len(seqs)
            thr = (1.0, 0.5, 0.0)[it % 3]
            if it % 2:
                cols, vals, nh, U = ix.search_sequence(seqs[j].encode(), k, h, thr)
            else:
                cols, vals, nh, U = ix.search_sequence_wait(ix.search_sequence_submit(seqs[j].encode(), k, h, thr))
            eU, cnt = want[j]
            if eU == 0:  # shorter than k: the C ABI reports no window (U = 0, no hits); BIGSI.search raises like the reference
                bad += not (U == 0 and nh == 0)
                continue
            exp = np.nonzero(cnt >= max(math.ceil(eU * thr), 0))[0]
            ok = U == eU and nh == len(exp) and np.array_equal(cols, exp) and np.array_equal(vals, cnt[exp])
            if not ok:
                bad += 1
                if bad <= 12:
                    print("MISMATCH k=%d N=%d it=%d seq=%d thr=%.1f sync=%d: U=%d (0x%x) want %d; n_hits=%d want %d; cols=%s vals=%s want vals=%s"
                          % (k, N, it, j, thr, it % 2, U, U, eU, nh, len(exp), cols[:8], vals[:8], cnt[exp][:8]))
        ix.close()
    print("seq_stress: %d mismatches" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
