"""A storage engine for the REFERENCE package whose bit-matrix rows live in HBM.

The reference picks its row store through STORAGE_DICT / get_storage(config) (bigsi/storage/__init__.py:1-19) and
talks to it through BaseStorage (bigsi/storage/base.py:9-151).  `register()` adds the engine "b200" to that table:
every "<row>:bitarray" value with an integer row key is kept in a bigsi_b200.DeviceIndex (upload_rows /
download_rows; a column insert -- BitMatrix.insert_column's set_bits over all rows, matrix/bitmatrix.py:67-75 -- is
ONE set_column kernel instead of m read-modify-writes), everything else (integers, strings, rows under
non-integer keys) in a host dictionary, as the reference's own engines keep them next to the rows.

Purpose: the reference's own test-suite (bigsi/tests/{bloom,graph,matrix,storage}) runs unmodified on top of the
HBM row store by appending a "b200" config to bigsi.tests.base.CONFIGS (oracle/run_reference_tests.py --engine b200,
tests/test_reference_injection.py), and a reference installation can keep its Python search path while its rows sit
on the GPU.  The fast path is bigsi_b200.BIGSI, not this adapter: here the reference still ANDs rows in Python.
"""
import numpy as np

from .index import DeviceIndex

_STORES = {}  # name -> _HbmRows: a named store persists for the life of the process, like a DB file


class _HbmRows:
    """Mapping bytes -> bytes.  Keys b"<int>:bitarray" -> rows of a DeviceIndex (grown on demand), others -> dict."""

    def __init__(self, device=0, index_factory=None):
        self.device = device
        self._make = index_factory or (lambda rows, cols, dev: DeviceIndex(rows, cols, device=dev))  # (tests inject a host stand-in)
        self.other = {}
        self.index = None
        self.rows_cap = 0       # rows the DeviceIndex holds
        self.bytes_cap = 0      # bytes per row it holds
        self.row_len = {}       # row id -> logical length in bytes (the reference's rows are as long as they were set)

    # -- helpers ----------------------------------------------------------------------------------------------
    @staticmethod
    def _row_of(key):
        if key.endswith(b":bitarray"):
            head = key[:-9]
            if head.isdigit():
                return int(head)
        return None

    def _ensure(self, row, nbytes):
        if self.index is not None and row < self.rows_cap and nbytes <= self.bytes_cap:
            return
        rows_cap = max(self.rows_cap, 64)
        while rows_cap <= row:
            rows_cap *= 2
        bytes_cap = max(self.bytes_cap, 16)
        while bytes_cap < nbytes:
            bytes_cap *= 2
        new = self._make(rows_cap, bytes_cap * 8, self.device)
        if self.index is not None:
            if self.row_len:
                top = max(self.row_len) + 1
                old = self.index.download_rows(0, top)
                pad = np.zeros((top, bytes_cap), dtype=np.uint8)
                pad[:, : old.shape[1]] = old
                new.upload_rows(0, pad)
            self.index.close()
        self.index, self.rows_cap, self.bytes_cap = new, rows_cap, bytes_cap

    # -- mapping protocol -------------------------------------------------------------------------------------
    def __setitem__(self, key, val):
        row = self._row_of(key)
        if row is None:
            self.other[key] = val
            return
        val = bytes(val)
        self._ensure(row, len(val))
        buf = np.zeros((1, self.bytes_cap), dtype=np.uint8)  # full width: the tail of a shorter row is zero in HBM
        buf[0, : len(val)] = np.frombuffer(val, dtype=np.uint8)
        self.index.upload_rows(row, buf)
        self.row_len[row] = len(val)

    def __getitem__(self, key):
        row = self._row_of(key)
        if row is None:
            return self.other[key]
        if row not in self.row_len:
            raise KeyError(key)
        return self.index.download_rows(row, 1)[0, : self.row_len[row]].tobytes()

    def get_rows(self, rows):
        """bytes of several rows with ONE download per contiguous run."""
        out = {}
        rows = list(rows)
        for r in rows:
            if r not in self.row_len:
                raise KeyError(b"%d:bitarray" % r)
        for r in sorted(set(rows)):
            if r in out:
                continue
            stop = r
            while stop + 1 in self.row_len and stop + 1 - r < 4096:
                stop += 1
            block = self.index.download_rows(r, stop - r + 1)
            for i in range(r, stop + 1):
                out[i] = block[i - r, : self.row_len[i]].tobytes()
        return [out[r] for r in rows]

    def set_column(self, n_rows, col, bits):
        """Column `col` of rows [0, n_rows) <- bits (sequence of 0/1): one kernel.  Rows grow by one byte when col
        is their next free bit (BaseStorage.set_bit appends, storage/base.py:118-124)."""
        need = col // 8 + 1
        self._ensure(n_rows - 1, need)
        packed = np.packbits(np.asarray(bits, dtype=np.uint8))
        self.index.set_column(col, packed, n_rows)
        for r in range(n_rows):
            if self.row_len.get(r, 0) < need:
                self.row_len[r] = need

    def clear(self):
        self.other.clear()
        self.row_len.clear()
        if self.index is not None:
            self.index.close()
        self.index, self.rows_cap, self.bytes_cap = None, 0, 0


def make_storage_class(BaseStorage):
    """B200Storage bound to the BaseStorage of the reference package that is imported in this process."""

    class B200Storage(BaseStorage):
        def __init__(self, storage_config=None):
            cfg = storage_config or {}
            self.name = cfg.get("filename", "default")
            key = (self.name, int(cfg.get("device", 0)))
            self.storage = _STORES.setdefault(key, _HbmRows(device=key[1]))

        def batch_get(self, keys):  # BitMatrix.get_rows -> get_bitarrays -> batch_get (storage/base.py:96-109)
            keys = [k if isinstance(k, bytes) else self.convert_key_to_bytes(k) for k in keys]
            rows = [self.storage._row_of(k) for k in keys]
            if keys and all(r is not None for r in rows):
                return self.storage.get_rows(rows)
            return [self[k] for k in keys]

        def set_bits(self, keys, positions, bits):  # BitMatrix.insert_column (matrix/bitmatrix.py:67-75)
            keys, positions, bits = list(keys), list(positions), list(bits)
            n = len(keys)
            col = positions[0] if positions else 0
            rows_ok = all(isinstance(k, int) for k in keys) and keys == list(range(n)) and all(p == col for p in positions)
            st = self.storage
            if n and rows_ok and all(r in st.row_len and (col < 8 * st.row_len[r] or col == 8 * st.row_len[r]) for r in range(n)):
                st.set_column(n, col, [1 if b else 0 for b in bits])
                return
            BaseStorage.set_bits(self, keys, positions, bits)

        def delete_all(self):
            self.storage.clear()

        def close(self):
            pass

    return B200Storage


def register():
    """STORAGE_DICT["b200"] = B200Storage in the reference package importable in this process; returns the class."""
    from bigsi.storage import STORAGE_DICT
    from bigsi.storage.base import BaseStorage

    if "b200" not in STORAGE_DICT:
        STORAGE_DICT["b200"] = make_storage_class(BaseStorage)
    return STORAGE_DICT["b200"]


def b200_config(name, k, m, h, device=0):
    return {"storage-engine": "b200", "storage-config": {"filename": name, "device": device}, "k": k, "m": m, "h": h}
