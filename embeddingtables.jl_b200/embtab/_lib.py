"""ctypes binding of libembtab_b200.so (include/embtab_b200.h).

This is the ONLY way the host mirror computes anything: there is no CPU fallback.  If the
shared library is missing, or a call is made without a CUDA device, the error is raised to the
caller (EmbTabError) -- never papered over.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ETB_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libembtab_b200.so")

F32, F64, I32, I64, F16, BF16 = 0, 1, 2, 3, 4, 5
UPDATE_FMA, UPDATE_SPLIT_LONG = 1, 2


class EmbTabError(RuntimeError):
    pass


class Table(C.Structure):
    _fields_ = [("base", C.c_void_p), ("chunks", C.c_void_p), ("nrows", C.c_int64),
                ("shard_rows", C.c_int64), ("dim", C.c_int32), ("ld", C.c_int32),
                ("elt", C.c_int32), ("reserved", C.c_int32)]


class LookupItem(C.Structure):
    _fields_ = [("table", Table), ("idx", C.c_void_p), ("dst", C.c_void_p), ("ld_dst", C.c_int64),
                ("batch", C.c_int64), ("bag", C.c_int64), ("ld_idx", C.c_int64),
                ("idx_elt", C.c_int32), ("reserved", C.c_int32)]


class UpdateItem(C.Structure):
    _fields_ = [("table", Table), ("delta", C.c_void_p), ("ld_delta", C.c_int64),
                ("idx", C.c_void_p), ("batch", C.c_int64), ("bag", C.c_int64),
                ("ld_idx", C.c_int64), ("idx_elt", C.c_int32), ("flags", C.c_int32)]


class IndexView(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("map", C.c_void_p), ("records", C.c_void_p),
                ("nnz", C.c_void_p), ("scratch", C.c_void_p), ("n_total", C.c_int64),
                ("key_bytes", C.c_int32), ("row_bits", C.c_int32), ("num_splits", C.c_int32),
                ("this_split", C.c_int32)]


_SIGS = {
    "etb_version": ([], C.c_int32),
    "etb_last_error": ([], C.c_char_p),
    "etb_last_launch_count": ([], C.c_int32),
    "etb_device_count": ([C.POINTER(C.c_int32)], C.c_int32),
    "etb_init": ([C.c_int32], C.c_int32),
    "etb_malloc": ([C.POINTER(C.c_void_p), C.c_size_t], C.c_int32),
    "etb_free": ([C.c_void_p], C.c_int32),
    "etb_malloc_host": ([C.POINTER(C.c_void_p), C.c_size_t], C.c_int32),
    "etb_free_host": ([C.c_void_p], C.c_int32),
    "etb_memcpy_h2d": ([C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p], C.c_int32),
    "etb_memcpy_d2h": ([C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p], C.c_int32),
    "etb_memcpy_d2d": ([C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p], C.c_int32),
    "etb_memcpy2d_h2d": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p], C.c_int32),
    "etb_memcpy2d_d2h": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p], C.c_int32),
    "etb_memcpy2d_d2d": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p], C.c_int32),
    "etb_memset": ([C.c_void_p, C.c_int32, C.c_size_t, C.c_void_p], C.c_int32),
    "etb_stream_create": ([C.POINTER(C.c_void_p)], C.c_int32),
    "etb_stream_sync": ([C.c_void_p], C.c_int32),
    "etb_stream_destroy": ([C.c_void_p], C.c_int32),
    "etb_gather": ([C.c_void_p, C.c_int64, C.POINTER(Table), C.c_void_p, C.c_int32, C.c_int64,
                    C.c_void_p], C.c_int32),
    "etb_pooled_sum": ([C.c_void_p, C.c_int64, C.POINTER(Table), C.c_void_p, C.c_int32, C.c_int64,
                        C.c_int64, C.c_int64, C.c_void_p], C.c_int32),
    "etb_maplookup": ([C.POINTER(LookupItem), C.c_int32, C.c_void_p], C.c_int32),
    "etb_index_workspace_bytes": ([C.POINTER(UpdateItem), C.c_int32, C.POINTER(C.c_size_t)], C.c_int32),
    "etb_index": ([C.c_void_p, C.c_size_t, C.POINTER(UpdateItem), C.c_int32, C.POINTER(IndexView),
                   C.c_void_p], C.c_int32),
    "etb_sgd_update": ([C.POINTER(IndexView), C.POINTER(UpdateItem), C.c_int32, C.c_double, C.c_int32,
                        C.c_void_p], C.c_int32),
    "etb_adagrad_update": ([C.POINTER(IndexView), C.POINTER(UpdateItem), C.POINTER(C.c_void_p), C.c_int32, C.c_double,
                            C.c_double, C.c_int32, C.c_void_p], C.c_int32),
    "etb_index_and_update": ([C.c_void_p, C.c_size_t, C.POINTER(UpdateItem), C.c_int32, C.c_double,
                              C.c_int32, C.c_void_p], C.c_int32),
    "etb_uncompress": ([C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                        C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p], C.c_int32),
    "etb_a2a_unpack": ([C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                        C.c_int32, C.c_int64, C.c_int32, C.c_void_p], C.c_int32),
    "etb_ipc_export": ([C.c_void_p, C.c_void_p], C.c_int32),
    "etb_ipc_import": ([C.c_void_p, C.POINTER(C.c_void_p)], C.c_int32),
    "etb_ipc_close": ([C.c_void_p], C.c_int32),
    "etb_a2a_scatter": ([C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                         C.c_int32, C.c_int64, C.c_int32, C.c_void_p], C.c_int32),
    "etb_a2a_scatter_ld": ([C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                            C.POINTER(C.c_int64), C.c_int32, C.c_int64, C.c_int32, C.c_void_p], C.c_int32),
    "etb_peer_barrier": ([C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_uint32, C.c_void_p], C.c_int32),
    "etb_cache_admit": ([C.POINTER(IndexView), C.POINTER(UpdateItem), C.c_int32, C.c_int32, C.c_void_p], C.c_int32),
    "etb_cache_flush": ([C.POINTER(Table), C.c_void_p], C.c_int32),
    "etb_a2a_pack": ([C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                      C.c_int32, C.c_int64, C.c_int32, C.c_void_p], C.c_int32),
}

_lib = None


def lib():
    """The loaded shared library.  Raises EmbTabError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EmbTabError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C embeddingtables.jl_b200/csrc`).  There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(l, name)
            fn.argtypes, fn.restype = args, res
        _lib = l
    return _lib


def check(status: int):
    if status != 0:
        raise EmbTabError(f"libembtab_b200 status {status}: {lib().etb_last_error().decode()}")


def exported_symbols():
    return sorted(_SIGS)
