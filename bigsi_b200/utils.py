"""Host-side helpers of the query path (mirror of bigsi/utils/fncts.py in the reference).

These keep the reference's exact semantics: k-mers are every window of
length k, canonical = ASCII-lexicographic min of a k-mer and its reverse complement with only
A<->T, C<->G complemented.
"""
from functools import reduce

import numpy as np

_COMPLEMENT = str.maketrans("ACGT", "TGCA")


def seq_to_kmers(seq, kmer_size):
    """bigsi/utils/fncts.py:63-65."""
    for i in range(len(seq) - kmer_size + 1):
        yield seq[i : i + kmer_size]


def reverse_comp(s):
    """bigsi/utils/fncts.py:38-39 (bases outside ACGT pass through unchanged)."""
    return s.translate(_COMPLEMENT)[::-1]


def canonical(k):
    """bigsi/utils/fncts.py:51-54."""
    r = reverse_comp(k)
    return k if k <= r else r


def convert_query_kmer(kmer):
    return canonical(kmer)


def convert_query_kmers(kmers):
    for k in kmers:
        yield convert_query_kmer(k)


def bitwise_and(bitarrays):
    """bigsi/utils/fncts.py:24-25 (TypeError on an empty sequence, like the reference)."""
    return reduce(lambda x, y: x & y, bitarrays)


def non_zero_bitarrary_positions(bits):
    """bigsi/utils/fncts.py:28-29: positions of the set bits, ascending."""
    return np.nonzero(np.asarray(bits, dtype=bool))[0].tolist()


def chunks(l, n):
    for i in range(0, len(l), n):
        yield l[i : i + n]


def unique_kmers(kmers):
    """set(kmers) of bigsi/graph/index.py:45 with a deterministic first-occurrence order."""
    return list(dict.fromkeys(kmers))
