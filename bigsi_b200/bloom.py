"""Bloom-filter primitives (mirror of bigsi/bloom/bloomfilter.py); hashing runs on the GPU."""
import numpy as np

from . import bits as _bits
from .index import bloom_kmers, hash_kmers


def generate_hashes(element, number_hash_functions, bloomfilter_size, device=0):
    """bigsi/bloom/bloomfilter.py:9-13: {mmh3.hash(element, seed) % m for seed in range(h)} --
    signed MurmurHash3_x86_32 with Python floor-mod, NO canonicalisation at this level."""
    b = element.encode("utf-8") if isinstance(element, str) else bytes(element)
    arr = np.frombuffer(b, dtype=np.uint8).reshape(1, len(b))
    r = hash_kmers(arr, len(b), number_hash_functions, bloomfilter_size, canonical=False, device=device)
    return {int(x) for x in r.reshape(-1)}


class BloomFilter(object):
    """bigsi/bloom/bloomfilter.py:16-32.  The filter is kept as the packed MSB-first bytes of the
    reference's bitarray; `update` hashes the elements AND sets the bits on the GPU
    (bigsi_b200_bloom_kmers).  Bits are zero-initialised here (the reference leaves `bitarray(m)`
    uninitialised, which is why its shipped .bloom fixtures carry stray bits)."""

    def __init__(self, m, h, device=0):
        self.m = m
        self.h = h
        self.device = device
        self._packed = np.zeros((m + 7) // 8, dtype=np.uint8)

    @property
    def bitarray(self):
        return _bits.from_packed(self._packed, self.m)

    def add(self, e):
        for i in generate_hashes(e, self.h, self.m, self.device):
            self._packed[i >> 3] |= 0x80 >> (i & 7)

    def update(self, elements):
        elements = list(elements)
        if not elements:
            return self
        lens = {len(e) for e in elements}
        if len(lens) == 1:
            k = lens.pop()
            arr = np.frombuffer("".join(elements).encode("utf-8"), dtype=np.uint8)
            if k and arr.size == len(elements) * k:
                self._packed |= bloom_kmers(arr.reshape(len(elements), k), k, self.h, self.m, canonical=False,
                                            device=self.device)
                return self
        for e in elements:
            self.add(e)
        return self


def load_bitarray(f):
    """bigsi/bloom/bloomfilter.py:35-39: a .bloom file is the raw MSB-first bytes of the filter."""
    with open(f, "rb") as inf:
        data = np.frombuffer(inf.read(), dtype=np.uint8)
    return _bits.from_packed(data, data.size * 8)
