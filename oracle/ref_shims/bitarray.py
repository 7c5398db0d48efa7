"""TEST INFRASTRUCTURE ONLY -- numpy-backed stand-in for the third-party C
`bitarray` package (unpinned, /root/reference/requirements.txt:5), implementing the
subset the reference uses (storage/base.py:86-122, utils/fncts.py:24-29,
graph/bigsi.py:39-53, matrix/transpose.py:14-43, bloom/bloomfilter.py:16-39).
Big-endian bit order (bitarray default): bit i lives in byte i>>3, mask 0x80>>(i&7);
tobytes() zero-pads the last byte."""
import numpy as np


class bitarray:
    __slots__ = ("_b", "_n")

    def __init__(self, init=None, endian="big"):
        assert endian == "big"
        if init is None:
            self._b = np.zeros(0, dtype=np.uint8)
            self._n = 0
        elif isinstance(init, (int, np.integer)):
            # real bitarray(n) is uninitialised memory; zeros is one legal outcome
            self._n = int(init)
            self._b = np.zeros((self._n + 7) // 8, dtype=np.uint8)
        elif isinstance(init, str):
            bits = np.frombuffer(init.encode("ascii"), dtype=np.uint8) - ord("0")
            assert ((bits == 0) | (bits == 1)).all()
            self._n = len(bits)
            self._b = np.packbits(bits.astype(np.uint8))
        elif isinstance(init, bitarray):
            self._n = init._n
            self._b = init._b.copy()
        else:
            bits = np.array([1 if x else 0 for x in init], dtype=np.uint8)
            self._n = len(bits)
            self._b = np.packbits(bits)

    # -- helpers -----------------------------------------------------
    @classmethod
    def _from(cls, b, n):
        o = cls.__new__(cls)
        o._b = b
        o._n = n
        return o

    def _bits(self):
        return np.unpackbits(self._b)[: self._n]

    # -- container protocol -----------------------------------------
    def __len__(self):
        return self._n

    def length(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            start, stop, step = i.indices(self._n)
            if step == 1 and start == 0:
                n = max(0, stop)
                nb = (n + 7) // 8
                b = self._b[:nb].copy()
                if n & 7:
                    b[-1] &= (0xFF << (8 - (n & 7))) & 0xFF
                return bitarray._from(b, n)
            bits = self._bits()[i]
            return bitarray._from(np.packbits(bits), len(bits))
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("bitarray index out of range")
        return bool((self._b[i >> 3] >> (7 - (i & 7))) & 1)

    def __setitem__(self, i, v):
        if isinstance(i, slice):
            bits = self._bits().copy()
            bits[i] = np.array([1 if x else 0 for x in v], dtype=np.uint8) if not isinstance(v, (bool, int)) else (1 if v else 0)
            self._b = np.packbits(bits)
            return
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("bitarray assignment index out of range")
        m = 0x80 >> (i & 7)
        if v:
            self._b[i >> 3] |= m
        else:
            self._b[i >> 3] &= (~m) & 0xFF

    def __iter__(self):
        return (bool(x) for x in self._bits())

    def __eq__(self, o):
        return isinstance(o, bitarray) and self._n == o._n and bool((self._norm() == o._norm()).all())

    def __ne__(self, o):
        return not self.__eq__(o)

    __hash__ = None

    def _norm(self):
        b = self._b.copy()
        if self._n & 7:
            b[-1] &= (0xFF << (8 - (self._n & 7))) & 0xFF
        return b

    def __and__(self, o):
        if self._n != o._n:
            raise ValueError("bitarrays of equal length expected")
        return bitarray._from(np.bitwise_and(self._b, o._b), self._n)

    def __or__(self, o):
        if self._n != o._n:
            raise ValueError("bitarrays of equal length expected")
        return bitarray._from(np.bitwise_or(self._b, o._b), self._n)

    def __repr__(self):
        return "bitarray('%s')" % self.to01()

    def __array__(self, dtype=None, copy=None):
        a = self._bits().astype(bool)
        return a if dtype is None else a.astype(dtype)

    # -- bitarray API -----------------------------------------------
    def append(self, v):
        self._n += 1
        if (self._n + 7) // 8 > len(self._b):
            self._b = np.concatenate([self._b, np.zeros(1, dtype=np.uint8)])
        self[self._n - 1] = v

    def extend(self, it):
        if isinstance(it, bitarray):
            bits = np.concatenate([self._bits(), it._bits()])
        else:
            if isinstance(it, str):
                it = [c == "1" for c in it]
            bits = np.concatenate([self._bits(), np.array([1 if x else 0 for x in it], dtype=np.uint8)])
        self._n = len(bits)
        self._b = np.packbits(bits)

    def setall(self, v):
        self._b[:] = 0xFF if v else 0
        if v and self._n & 7:
            self._b[-1] &= (0xFF << (8 - (self._n & 7))) & 0xFF

    def count(self, v=True):
        c = int(self._bits().sum())
        return c if v else self._n - c

    def any(self):
        return bool(self._norm().any())

    def tolist(self):
        return [bool(x) for x in self._bits()]

    def to01(self):
        return "".join("1" if x else "0" for x in self._bits())

    def tobytes(self):
        return self._norm().tobytes()

    def frombytes(self, data):
        add = np.frombuffer(bytes(data), dtype=np.uint8)
        if self._n & 7:
            bits = np.concatenate([self._bits(), np.unpackbits(add)])
            self._n = len(bits)
            self._b = np.packbits(bits)
        else:
            self._b = np.concatenate([self._b[: self._n // 8], add])
            self._n += 8 * len(add)

    def tofile(self, f):
        f.write(self.tobytes())

    def fromfile(self, f, n=-1):
        self.frombytes(f.read() if n < 0 else f.read(n))

    def unpack(self, zero=b"\x00", one=b"\xff"):
        z, o = zero[0], one[0]
        bits = self._bits()
        return np.where(bits.astype(bool), np.uint8(o), np.uint8(z)).astype(np.uint8).tobytes()

    def pack(self, data):
        bits = (np.frombuffer(bytes(data), dtype=np.uint8) != 0).astype(np.uint8)
        self.extend(bitarray._from(np.packbits(bits), len(bits)))

    def copy(self):
        return bitarray._from(self._b.copy(), self._n)
