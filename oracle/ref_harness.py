"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference package from
/root/reference (read-only, only present in the build container) through the
stand-ins in oracle/ref_shims, so golden vectors can be generated from the real
reference code (tests/golden/make_golden.py) and the restatements in oracle/ can
be cross-checked against it.  Never imported by bigsi_b200/.  On the GPU box the
tree is absent; the unmodified package installed into the git-ignored baseline/_ref/
(oracle/install_reference.py) is imported instead -- by the reference-suite injection
test and by the "reference" CPU timing of bench.py only.

Recipe:
  * `mmh3`, `bitarray`, `redis` stand-ins ahead of the reference on sys.path;
  * `np.fromstring` (removed for binary input in NumPy 2) mapped to
    `np.frombuffer` for /root/reference/bigsi/graph/bigsi.py:39-53;
  * a dict-backed BaseStorage subclass registered as STORAGE_DICT["dict"].
"""
import os
import sys

_INSTALLED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("BIGSI_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "bigsi")) and os.path.isdir(os.path.join(_INSTALLED, "bigsi")):
    # the GPU box: no /root/reference, but the unmodified package installed by oracle/install_reference.py travels
    REFERENCE_ROOT = _INSTALLED
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")

_DICT_STORES = {}


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bigsi"))


def load_reference():
    """Returns the imported reference `bigsi` module (with a "dict" storage engine)."""
    if not reference_available():
        raise RuntimeError("reference tree %s not present" % REFERENCE_ROOT)
    import numpy as np

    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)

    _orig_fromstring = getattr(np, "fromstring", None)

    def _fromstring(s, dtype=float, count=-1, sep=""):
        if sep == "" and isinstance(s, (bytes, bytearray, memoryview)):
            return np.frombuffer(s, dtype=dtype, count=count).copy()
        return _orig_fromstring(s, dtype=dtype, count=count, sep=sep)

    np.fromstring = _fromstring

    import logging

    logging.disable(logging.WARNING)
    import bigsi  # the reference package, unmodified
    from bigsi.storage import STORAGE_DICT
    from bigsi.storage.base import BaseStorage

    if "dict" not in STORAGE_DICT:

        class DictStorage(BaseStorage):
            """In-memory KV store with the BaseStorage contract
            (/root/reference/bigsi/storage/base.py:9-151); named stores persist for the
            life of the process like a DB file would."""

            def __init__(self, storage_config=None):
                name = (storage_config or {}).get("filename", "default")
                self.name = name
                self.storage = _DICT_STORES.setdefault(name, {})

            def delete_all(self):
                _DICT_STORES[self.name].clear()
                self.storage = _DICT_STORES[self.name]

            def close(self):
                pass

        STORAGE_DICT["dict"] = DictStorage
    return bigsi


def dict_config(name, k, m, h):
    return {"storage-engine": "dict", "storage-config": {"filename": name}, "k": k, "m": m, "h": h}
