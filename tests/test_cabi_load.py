"""CPU: the C-ABI library loads and exports every symbol include/bigsi_b200.h declares; argument
validation that needs no device works; compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bigsi_b200.build import build

    build()
    from bigsi_b200 import _lib

    return _lib


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "bigsi_b200.h")).read()
    declared = set(re.findall(r"\b(bigsi_b200_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), "missing export %s" % name
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert L.bigsi_b200_abi_version() == int(re.search(r"BIGSI_B200_ABI_VERSION (\d+)", hdr).group(1))


def test_info_struct_matches_header(lib):
    hdr = open(os.path.join(ROOT, "include", "bigsi_b200.h")).read()
    body = hdr[hdr.index("typedef struct {") : hdr.index("} bigsi_b200_info;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(u?int\d+_t)\s+([^;]+);", body):
        for name in decl[1].split(","):
            fields.append((name.strip(), decl[0]))
    assert [f for f, _ in fields] == [f for f, _ in lib.Info._fields_]
    ctype = {"uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32}
    assert [ctype[t] for _, t in fields] == [t for _, t in lib.Info._fields_]


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = lib.lib()
    h = ctypes.c_void_p(0)
    rc = L.bigsi_b200_index_create(0, 1000, 8, 0, 0, ctypes.byref(h))
    assert rc in (lib.ERR_NO_DEVICE, lib.ERR_CUDA) and not h.value
    assert L.bigsi_b200_last_error()
    import bigsi_b200

    with pytest.raises(bigsi_b200.BigsiB200Error):
        bigsi_b200.generate_hashes("ATT", 3, 25)
    with pytest.raises(bigsi_b200.BigsiB200Error):
        bigsi_b200.DeviceIndex(1000, 8)


def test_argument_validation_without_device(lib):
    L = lib.lib()
    assert L.bigsi_b200_index_create(0, 0, 8, 0, 0, ctypes.byref(ctypes.c_void_p(0))) == lib.ERR_INVALID
    assert L.bigsi_b200_index_create(0, 10, 8, 0, 3, ctypes.byref(ctypes.c_void_p(0))) == lib.ERR_INVALID
    assert L.bigsi_b200_index_get_info(None, None) == lib.ERR_INVALID
    assert L.bigsi_b200_index_destroy(None) == 0
    assert L.bigsi_b200_hash_kmers_dev(None, 5, 0, 3, 25, 1, None, None) == lib.ERR_INVALID
    assert L.bigsi_b200_hash_kmers_dev(None, 5, 31, 3, 0, 1, None, None) == lib.ERR_INVALID
