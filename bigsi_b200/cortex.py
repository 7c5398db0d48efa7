"""McCortex graph (.ctx, format version 6) -> k-mers: the ingest side of `bigsi bloom`
(bigsi/cmds/bloom.py:17-27 -> bigsi/utils/cortex.py:23-27, 170-264).

The reference walks the file record by record in Python (struct.unpack + per-character string work
per k-mer); here the payload is decoded with numpy in one pass: the first 8 bytes of every record are
the k-mer, two bits per base (A=0, C=1, G=2, T=3), the LAST base in the lowest bit pair; the record's
canonical form (ASCII-lexicographic minimum of the k-mer and its reverse complement,
utils/cortex.py:99-105) is what `extract_kmers_from_ctx` yields -- cut into windows when a shorter k is
asked for.  Coverages and edges are skipped.  `write_ctx` produces minimal files for tests and tools.
"""
import struct

import numpy as np

MAGIC = b"CORTEX"
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[list(b"ACGT")] = list(b"TGCA")


def read_header(buf):
    """(kmer_size, record_size, payload_offset) of a version-6 graph file held in `buf`
    (layout as parsed by utils/cortex.py:189-230)."""
    if buf[:6] != MAGIC:
        raise ValueError("File format mismatch")
    version, kmer_size, words, ncols = struct.unpack_from("<IIII", buf, 6)
    if version != 6:
        raise ValueError("File format version error; only 6 supported")
    off = 22 + 12 * ncols                       # per colour: mean read length (u32) + total sequence (u64)
    for _ in range(ncols):                      # sample names
        (n,) = struct.unpack_from("<I", buf, off)
        off += 4 + n
    off += 16 * ncols                           # error rates (long double)
    for _ in range(ncols):                      # cleaning counters + the name of the cleaning graph
        off += 12
        (n,) = struct.unpack_from("<I", buf, off)
        off += 4 + n
    if buf[off : off + 6] != MAGIC:
        raise ValueError("File format mismatch")
    return kmer_size, 8 * words + 5 * ncols, off + 6


def ctx_kmer_array(path):
    """uint8 [n_records, kmer_size]: the canonical k-mer of every record, in file order."""
    with open(path, "rb") as f:
        buf = f.read()
    k, rec, start = read_header(buf)
    assert k <= 31  # utils/cortex.py:39: one 64-bit word per k-mer
    n = (len(buf) - start) // rec
    if n == 0:
        return np.zeros((0, k), dtype=np.uint8)
    recs = np.frombuffer(buf, dtype=np.uint8, count=n * rec, offset=start).reshape(n, rec)
    words = np.ascontiguousarray(recs[:, :8]).view("<u8").reshape(n)
    shifts = (2 * (k - 1 - np.arange(k))).astype(np.uint64)
    fwd = _BASES[((words[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.intp)]
    rev = _COMP[fwd[:, ::-1]]
    # lexicographic comparison of the rows: the first differing position decides
    diff = fwd != rev
    first = diff.argmax(axis=1)
    idx = np.arange(n)
    use_rev = diff.any(axis=1) & (rev[idx, first] < fwd[idx, first])
    return np.where(use_rev[:, None], rev, fwd)


def extract_kmers_from_ctx(ctx, k):
    """utils/cortex.py:23-27: every window of length k of every record's canonical k-mer (the k-mer itself
    when k equals the graph's k-mer size), in file order, as strings."""
    arr = ctx_kmer_array(ctx)
    for row in arr:
        s = row.tobytes().decode("ascii")
        for i in range(len(s) - k + 1):
            yield s[i : i + k]


def write_ctx(path, kmers, kmer_size=None, sample_names=("sample",), coverage=1, edges=0):
    """A minimal version-6 graph with one 64-bit word per k-mer and len(sample_names) colours."""
    kmers = list(kmers)
    k = kmer_size if kmer_size is not None else len(kmers[0])
    ncols = len(sample_names)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    out = [MAGIC, struct.pack("<IIII", 6, k, 1, ncols)]
    out.append(b"".join(struct.pack("<IQ", 100, 1000) for _ in range(ncols)))
    for name in sample_names:
        b = name.encode("utf-8")
        out.append(struct.pack("<I", len(b)) + b)
    out.append(b"\0" * (16 * ncols))
    for _ in range(ncols):
        g = b"undefined"
        out.append(b"\0" * 12 + struct.pack("<I", len(g)) + g)
    out.append(MAGIC)
    for km in kmers:
        v = 0
        for j, c in enumerate(reversed(km)):
            v |= code[c] << (2 * j)
        out.append(struct.pack("<Q", v) + struct.pack("<" + "I" * ncols, *([coverage] * ncols)) + bytes([edges] * ncols))
    with open(path, "wb") as f:
        f.write(b"".join(out))
