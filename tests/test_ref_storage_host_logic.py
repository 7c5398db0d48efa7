"""Host logic of the reference-side storage adapter (bigsi_b200/ref_storage.py) on CPU: the row store behind
B200Storage with a numpy stand-in for the DeviceIndex -- variable-length rows, growth of both dimensions, the
one-kernel column insert with the reference's append semantics (storage/base.py:111-124), non-row keys -- against a
plain dict model of BaseStorage.  (On a GPU the same adapter carries the reference's own suite:
tests/test_reference_injection.py.)"""
import numpy as np

from bigsi_b200 import ref_storage


class FakeIndex:
    """upload_rows / download_rows / set_column / close of bigsi_b200.DeviceIndex on a numpy array."""

    live = 0

    def __init__(self, rows, cols, device):
        assert cols % 8 == 0
        self.a = np.zeros((rows, cols // 8), dtype=np.uint8)
        FakeIndex.live += 1

    def upload_rows(self, row0, rows):
        rows = np.asarray(rows, dtype=np.uint8)
        self.a[row0: row0 + rows.shape[0], : rows.shape[1]] = rows
        self.a[row0: row0 + rows.shape[0], rows.shape[1]:] = 0

    def download_rows(self, row0, n):
        return self.a[row0: row0 + n].copy()

    def set_column(self, col, packed, n_bits):
        bits = np.unpackbits(np.asarray(packed, dtype=np.uint8))[:n_bits]
        byte, mask = col // 8, np.uint8(0x80 >> (col % 8))
        keep = np.uint8(0xFF ^ int(mask))
        self.a[:n_bits, byte] = np.where(bits == 1, self.a[:n_bits, byte] | mask, self.a[:n_bits, byte] & keep)

    def close(self):
        FakeIndex.live -= 1


def _store():
    return ref_storage._HbmRows(device=0, index_factory=FakeIndex)


def test_rows_of_any_length_grow_the_store():
    rng = np.random.default_rng(0)
    st, model = _store(), {}
    for step in range(400):
        row = int(rng.integers(0, 700)) if step % 7 else int(rng.integers(0, 5000))
        val = rng.integers(0, 256, size=int(rng.integers(0, 90)), dtype=np.uint8).tobytes()
        key = b"%d:bitarray" % row
        st[key] = val
        model[key] = val
        if step % 11 == 0:
            st[b"name:%d" % step] = val
            model[b"name:%d" % step] = val
            st[b"x%d:bitarray" % step] = val  # a row under a non-integer key stays on the host
            model[b"x%d:bitarray" % step] = val
    for k, v in model.items():
        assert st[k] == v, k
    rows = [int(k.split(b":")[0]) for k in model if k.endswith(b":bitarray") and k[:1].isdigit()]
    assert st.get_rows(rows) == [model[b"%d:bitarray" % r] for r in rows]
    for missing in (b"99999:bitarray", b"nokey"):
        try:
            st[missing]
            raise AssertionError("expected KeyError")
        except KeyError:
            pass
    st.clear()
    assert FakeIndex.live == 0 and not st.row_len and not st.other


def test_column_insert_appends_like_set_bit():
    """BaseStorage.set_bit: pos inside the row's 8*len bits sets in place, pos == 8*len appends one bit (-> one more byte)."""
    rng = np.random.default_rng(1)
    st = _store()
    n, width = 300, 3  # rows of 3 bytes = 24 bits
    rows = rng.integers(0, 256, size=(n, width), dtype=np.uint8)
    for r in range(n):
        st[b"%d:bitarray" % r] = rows[r].tobytes()
    model = np.unpackbits(rows, axis=1)  # n x 24
    for col in (5, 23, 24, 25, 31, 32):   # 24 and 32 are appends (a new byte), the others in place
        bits = rng.integers(0, 2, size=n, dtype=np.uint8)
        st.set_column(n, col, bits)
        if col >= model.shape[1]:
            model = np.concatenate([model, np.zeros((n, 8 * (col // 8 + 1) - model.shape[1]), dtype=np.uint8)], axis=1)
        model[:, col] = bits
        for r in (0, 1, n // 2, n - 1):
            assert st[b"%d:bitarray" % r] == np.packbits(model[r]).tobytes(), (col, r)
    st.clear()
