"""Bit-vector container returned by BIGSI.bloom()/lookup().

The reference hands out third-party `bitarray` objects (bigsi/bloom/bloomfilter.py:16-32,
storage/base.py:86-99).  When that package is installed it is used as is; otherwise a minimal
numpy-backed stand-in with the same observable behaviour for the subset BIGSI exposes
(MSB-first tobytes/frombytes, ==, &, len, indexing, tolist, to01, count) is provided.
"""
import numpy as np

try:  # pragma: no cover - not installed in the build image
    from bitarray import bitarray as _real_bitarray
except ImportError:
    _real_bitarray = None


class _PackedBits:
    """Subset of bitarray.bitarray (big-endian bit order) over a numpy bool vector."""

    __slots__ = ("_b",)

    def __init__(self, init=0):
        if isinstance(init, (int, np.integer)):
            self._b = np.zeros(int(init), dtype=bool)
        elif isinstance(init, str):
            self._b = np.frombuffer(init.encode("ascii"), dtype=np.uint8) == ord("1")
            if not np.all((self._b) | (np.frombuffer(init.encode("ascii"), dtype=np.uint8) == ord("0"))):
                raise ValueError("bit string must contain only '0' and '1'")
            self._b = self._b.copy()
        elif isinstance(init, _PackedBits):
            self._b = init._b.copy()
        else:
            self._b = np.array([bool(x) for x in init], dtype=bool)

    @classmethod
    def _wrap(cls, bools):
        o = cls.__new__(cls)
        o._b = np.ascontiguousarray(bools, dtype=bool)
        return o

    # -- bytes ---------------------------------------------------------------
    def tobytes(self):
        return np.packbits(self._b).tobytes()

    def frombytes(self, data):
        self._b = np.concatenate([self._b, np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8)).astype(bool)])

    def tofile(self, f):
        f.write(self.tobytes())

    def fromfile(self, f):
        self.frombytes(f.read())

    # -- sequence ------------------------------------------------------------
    def __len__(self):
        return int(self._b.size)

    def length(self):
        return len(self)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return _PackedBits._wrap(self._b[i].copy())
        return bool(self._b[i])

    def __setitem__(self, i, v):
        self._b[i] = bool(v) if not isinstance(i, slice) else v

    def __iter__(self):
        return (bool(x) for x in self._b)

    def append(self, v):
        self._b = np.append(self._b, bool(v))

    def extend(self, other):
        self._b = np.concatenate([self._b, _as_bools(other)])

    def setall(self, v):
        self._b[:] = bool(v)

    def count(self, v=True):
        n = int(self._b.sum())
        return n if v else len(self) - n

    def tolist(self):
        return [bool(x) for x in self._b]

    def to01(self):
        return "".join("1" if x else "0" for x in self._b)

    def __and__(self, other):
        o = _as_bools(other)
        if o.size != self._b.size:
            raise ValueError("bitarrays of equal length expected")
        return _PackedBits._wrap(self._b & o)

    def __eq__(self, other):
        try:
            o = _as_bools(other)
        except TypeError:
            return NotImplemented
        return o.size == self._b.size and bool(np.array_equal(self._b, o))

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __array__(self, dtype=None, copy=None):
        return self._b.astype(dtype) if dtype is not None else self._b

    def __repr__(self):
        return "bitarray('%s')" % self.to01()


def _as_bools(x):
    if isinstance(x, _PackedBits):
        return x._b
    if _real_bitarray is not None and isinstance(x, _real_bitarray):
        return np.unpackbits(np.frombuffer(x.tobytes(), dtype=np.uint8))[: len(x)].astype(bool)
    if isinstance(x, str):
        return _PackedBits(x)._b
    if isinstance(x, np.ndarray):
        return x.astype(bool)
    if isinstance(x, (list, tuple)):
        return np.array([bool(v) for v in x], dtype=bool)
    raise TypeError("cannot interpret %r as a bit vector" % type(x))


bitarray = _real_bitarray if _real_bitarray is not None else _PackedBits


def from_packed(packed, nbits):
    """MSB-first packed uint8 -> bitarray of nbits."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    if _real_bitarray is not None:  # pragma: no cover
        b = _real_bitarray()
        b.frombytes(packed.tobytes())
        return b[:nbits]
    return _PackedBits._wrap(np.unpackbits(packed)[:nbits].astype(bool))


def to_packed(bits, nbits=None):
    """bitarray / bool vector / '0101' string / packed uint8 (with nbits) -> MSB-first packed uint8."""
    if isinstance(bits, np.ndarray) and bits.dtype == np.uint8 and nbits is not None and bits.size * 8 >= nbits \
            and bits.size == (nbits + 7) // 8:
        return np.ascontiguousarray(bits)
    if hasattr(bits, "tobytes") and not isinstance(bits, np.ndarray):
        return np.frombuffer(bits.tobytes(), dtype=np.uint8).copy()
    return np.packbits(_as_bools(bits))


def nbits_of(bits):
    if isinstance(bits, np.ndarray) and bits.dtype == np.uint8:
        return None
    return len(bits)
