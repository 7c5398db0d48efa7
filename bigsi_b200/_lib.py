"""ctypes binding of libbigsi_b200.so (include/bigsi_b200.h).

The shared library is the product; it is built in-tree by `bigsi_b200.build` /
`__graft_entry__.build()`.  Loading fails loudly if it is missing -- there is no CPU or
PyTorch fallback path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbigsi_b200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_NO_DEVICE, ERR_RANGE, ERR_TIMEOUT = -1, -2, -3, -4, -5, -6
MODE_COUNTS, MODE_AND = 0, 1

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)


class Info(ctypes.Structure):
    _fields_ = [
        ("num_rows", ctypes.c_uint64),
        ("num_cols", ctypes.c_uint64),
        ("col_capacity", ctypes.c_uint64),
        ("col_offset", ctypes.c_uint64),
        ("row_bytes", ctypes.c_uint64),
        ("row_pitch_bytes", ctypes.c_uint64),
        ("matrix_bytes", ctypes.c_uint64),
        ("device", ctypes.c_int32),
        ("sm_count", ctypes.c_int32),
        ("last_kmers", ctypes.c_uint64),
        ("last_algorithmic_bytes", ctypes.c_uint64),
        ("last_grid", ctypes.c_uint32),
        ("last_block", ctypes.c_uint32),
        ("last_smem_bytes", ctypes.c_uint32),
        ("last_tile_bytes", ctypes.c_uint32),
        ("last_n_tiles", ctypes.c_uint32),
        ("last_kmers_per_stage", ctypes.c_uint32),
        ("last_n_stages", ctypes.c_uint32),
        ("last_n_slices", ctypes.c_uint32),
        ("kernel_launches", ctypes.c_uint64),
        ("scratch_bytes", ctypes.c_uint64),
        ("last_fused", ctypes.c_uint32),
        ("last_reduce_grid", ctypes.c_uint32),
        ("last_unique_kmers", ctypes.c_uint64),
    ]

    def asdict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class FileHeader(ctypes.Structure):
    """bigsi_b200_file_header (include/bigsi_b200.h)."""

    _fields_ = [
        ("magic", ctypes.c_char * 8),
        ("version", ctypes.c_uint32),
        ("header_bytes", ctypes.c_uint32),
        ("num_rows", ctypes.c_uint64),
        ("num_cols", ctypes.c_uint64),
        ("col_offset", ctypes.c_uint64),
        ("row_bytes", ctypes.c_uint64),
        ("meta_bytes", ctypes.c_uint64),
        ("rows_offset", ctypes.c_uint64),
    ]


# name -> (restype, argtypes); every symbol include/bigsi_b200.h declares
_u64, _i64, _int, _vp = ctypes.c_uint64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
SIGNATURES = {
    "bigsi_b200_abi_version": (_int, []),
    "bigsi_b200_last_error": (ctypes.c_char_p, []),
    "bigsi_b200_device_count": (_int, [ctypes.POINTER(_int)]),
    "bigsi_b200_host_alloc": (_int, [_u64, c_void_pp]),
    "bigsi_b200_host_free": (_int, [_vp]),
    "bigsi_b200_index_create": (_int, [_int, _u64, _u64, _u64, _u64, c_void_pp]),
    "bigsi_b200_index_destroy": (_int, [_vp]),
    "bigsi_b200_index_get_info": (_int, [_vp, ctypes.POINTER(Info)]),
    "bigsi_b200_index_set_option": (_int, [_vp, ctypes.c_char_p, _i64]),
    "bigsi_b200_index_timing_collect": (_int, [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                               ctypes.POINTER(_u64)]),
    "bigsi_b200_index_debug_read": (_int, [_vp, _vp, _u64]),
    "bigsi_b200_index_status": (_int, [_vp]),
    "bigsi_b200_index_upload_rows": (_int, [_vp, _u64, _u64, _vp, _u64, _u64]),
    "bigsi_b200_index_download_rows": (_int, [_vp, _u64, _u64, _vp, _u64]),
    "bigsi_b200_index_set_column": (_int, [_vp, _u64, _vp, _u64]),
    "bigsi_b200_index_fill_synthetic": (_int, [_vp, _u64, _int, _vp, _vp, _int]),
    "bigsi_b200_hash_kmers": (_int, [_int, _vp, _u64, _int, _int, _u64, _int, _vp]),
    "bigsi_b200_hash_kmers_dev": (_int, [_vp, _u64, _int, _int, _u64, _int, _vp, _vp]),
    "bigsi_b200_query_dev": (_int, [_vp, _int, _vp, _vp, _u64, _u64, _u64, _int, _vp, _u64, _vp]),
    "bigsi_b200_query_hits_dev": (_int, [_vp, _vp, _vp, _u64, _u64, _u64, _int, _vp, _vp, _vp, _u64, _vp, _vp, _u64, _vp]),
    "bigsi_b200_query_kmers_hits_dev": (_int, [_vp, _vp, _int, _vp, _u64, _u64, _u64, _int, _vp, _vp, _vp, _u64, _vp, _vp,
                                               _u64, _vp]),
    "bigsi_b200_query_kmers_hits_stream_dev": (_int, [_vp, _vp, _int, _u64, _int, ctypes.c_uint32, _vp, _vp, _u64, _vp, _vp]),
    "bigsi_b200_index_flush": (_int, [_vp]),
    "bigsi_b200_lookup_dev": (_int, [_vp, _vp, _u64, _int, _vp, _u64, _vp]),
    "bigsi_b200_threshold_dev": (_int, [_vp, _u64, _u64, _u64, _vp, _vp, _vp, _u64, _vp, _vp]),
    "bigsi_b200_search_kmers": (_int, [_vp, _int, _vp, _vp, _u64, _int, _int, _vp, _u64]),
    "bigsi_b200_search_rows": (_int, [_vp, _int, _vp, _vp, _u64, _int, _vp, _u64]),
    "bigsi_b200_search_kmers_hits": (_int, [_vp, _vp, _vp, _u64, _int, _int, _vp, _vp, _vp, _u64, _vp]),
    "bigsi_b200_lookup_kmers": (_int, [_vp, _vp, _u64, _int, _int, _vp, _u64]),
    "bigsi_b200_search_sequence": (_int, [_vp, _vp, _u64, _int, _int, ctypes.c_double, _vp, _vp, _u64, _vp, _vp]),
    "bigsi_b200_search_sequence_submit": (_int, [_vp, _vp, _u64, _int, _int, ctypes.c_double, _u64, ctypes.POINTER(_u64)]),
    "bigsi_b200_search_sequence_wait": (_int, [_vp, _u64, _vp, _vp, _u64, _vp, _vp]),
    "bigsi_b200_search_sequences": (_int, [_vp, _vp, _vp, _u64, _int, _int, ctypes.c_double, _vp, _vp, _u64, _vp, _vp]),
    "bigsi_b200_bloom_kmers": (_int, [_int, _vp, _u64, _int, _int, _u64, _int, _vp]),
    "bigsi_b200_index_build_columns": (_int, [_vp, _u64, _u64, _vp, _u64, _u64]),
    "bigsi_b200_index_build_columns_dev": (_int, [_vp, _u64, _u64, _vp, _u64, _u64, _vp]),
    "bigsi_b200_sequence_presence": (_int, [_vp, _vp, _u64, _int, _int, _vp, _u64, _vp]),
    "bigsi_b200_index_save": (_int, [_vp, ctypes.c_char_p, _vp, _u64]),
    "bigsi_b200_file_info": (_int, [ctypes.c_char_p, _vp, _vp, _u64]),
    "bigsi_b200_index_load_rows": (_int, [_vp, ctypes.c_char_p, _u64, _u64, _u64, _u64, _u64]),
    "bigsi_b200_exchange_create": (_int, [_vp, _int, _int, _u64, ctypes.c_uint32, _vp]),
    "bigsi_b200_exchange_open": (_int, [_vp, _vp]),
    "bigsi_b200_exchange_open_local": (_int, [_vp, _vp]),
    "bigsi_b200_exchange_search_dev": (_int, [_vp, _vp, _u64, _int, _int, ctypes.c_uint32, _vp, c_void_pp,
                                              ctypes.POINTER(_u64)]),
    "bigsi_b200_exchange_reserve": (_int, [_vp, _u64, _int, _int]),
    "bigsi_b200_exchange_host_results": (_int, [_vp, _int]),
    "bigsi_b200_exchange_last_seq": (_int, [_vp, ctypes.POINTER(_u64)]),
    "bigsi_b200_exchange_wait_host": (_int, [_vp, _u64, c_void_pp, ctypes.POINTER(_u64)]),
    "bigsi_b200_exchange_wait_ns": (_int, [_vp, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "bigsi_b200_exchange_destroy": (_int, [_vp]),
}

_lib = None


class BigsiB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("bigsi_b200 error %d: %s" % (code, message))
        self.code = code


def lib():
    """The loaded library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build the CUDA extension first (python -m bigsi_b200.build). "
                "bigsi_b200 has no CPU fallback." % LIB_PATH
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = restype
            f.argtypes = argtypes
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        msg = lib().bigsi_b200_last_error()
        raise BigsiB200Error(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc


def device_count():
    n = ctypes.c_int(0)
    rc = lib().bigsi_b200_device_count(ctypes.byref(n))
    if rc != OK:
        return 0
    return n.value
