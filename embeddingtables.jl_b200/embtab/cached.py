"""CachedEmbedding: a host-tier table behind an HBM row cache (GPU-only extension).

The reference has no such table, but it leaves the hook for one: every kernel passes an IndexingContext --
Forward() from the lookups (src/lookup.jl:57-161), Update() from update! (src/sparseupdate.jl:25-122) -- to
`columnpointer(table, i, ctx)` (src/EmbeddingTables.jl:74-93), so that a table type may resolve a row differently
per phase.  Here the whole table lives in page-locked HOST memory (tables larger than the 180 GB of HBM), the GPU
addresses it directly over PCIe, and up to `cache_rows` of its rows have a copy in HBM:

    row address = slot_of_row[i-1] >= 0 ? cache + slot*ld : host_base + (i-1)*ld        (csrc/etb_common.cuh row_ptr)

  * Forward context: a pure read through that rule.
  * Update context: update! writes a cached row in HBM (it is authoritative there until `flush()`), an uncached row on
    the host; afterwards -- still the Update phase -- `etb_cache_admit` walks the bucket records that index! just
    produced: it histograms the occurrence counts of the rows still on the host and admits those at or above the
    count at which they fit into the free slots, never below `min_count` (the hottest rows of a Zipf batch come first;
    nothing is evicted, `flush()` + `clear()` start over).

Arithmetic does not depend on where a row lives: results equal the all-HBM tables' bit for bit
(tests/test_gpu_cached.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .darray import DeviceArray, current_stream_ptr, pinned_empty
from .tables import AbstractEmbeddingTable, ArgumentError, Dynamic, Forward, IndexingContext, Static, Update


class CacheDesc(C.Structure):   # etb_cache_desc (a host struct of device pointers)
    _fields_ = [("rows", C.c_void_p), ("slot_of_row", C.c_void_p), ("row_of_slot", C.c_void_p),
                ("cursor", C.c_void_p), ("capacity", C.c_int64), ("hist", C.c_void_p)]


TABLE_CACHED = -1   # ETB_TABLE_CACHED


class CachedEmbedding(AbstractEmbeddingTable):
    """CachedEmbedding(A, cache_rows, lookup_type=None, min_count=2): `A` (featuresize x nrows, host) is copied into
    page-locked memory; `cache_rows` HBM slots."""

    def __init__(self, A, cache_rows: int, lookup_type=None, min_count: int = 2):
        A = np.asarray(A)
        if A.ndim != 2:
            raise ArgumentError("CachedEmbedding wraps a matrix")
        if isinstance(lookup_type, Static) and lookup_type.N != A.shape[0]:
            raise ArgumentError(f"Parameter `N` should match the number of rows in the passed Matrix. Instead, `N = {lookup_type.N}` "
                                f"while `size(A,1) = {A.shape[0]}`.")
        self.lookup_type = lookup_type if lookup_type is not None else Dynamic()
        self.dtype = A.dtype
        self.host = pinned_empty(A.shape, A.dtype)            # the table itself: host memory the GPU can address
        self.host[...] = A
        self.capacity, self.min_count = int(cache_rows), int(min_count)
        f, n = A.shape
        self.cache = DeviceArray.empty((f, max(1, self.capacity)), A.dtype)
        self._slot_of_row = torch.full((max(1, n),), -1, dtype=torch.int32, device="cuda")
        self._row_of_slot = torch.zeros(max(1, self.capacity), dtype=torch.int32, device="cuda")
        self._cursor = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._hist = torch.zeros(64, dtype=torch.int32, device="cuda")      # ETB_CACHE_HIST_BINS
        self._desc = CacheDesc(self.cache.ptr, self._slot_of_row.data_ptr(), self._row_of_slot.data_ptr(),
                               self._cursor.data_ptr(), self.capacity, self._hist.data_ptr())
        self.context = None     # the IndexingContext of the last descriptor() request (introspection / tests)

    def size(self, d=None):
        s = self.host.shape
        return s if d is None else s[d - 1]

    def example(self) -> DeviceArray:
        return self.cache

    def columnpointer(self, i: int, ctx: IndexingContext = None) -> int:
        """where row i lives right now (host introspection; the kernels use the same rule on the device)"""
        s = int(self._slot_of_row[i - 1].item())
        stride = self.host.shape[0] * self.host.itemsize
        return self.cache.ptr + s * stride if s >= 0 else self.host.ctypes.data + (i - 1) * stride

    def descriptor(self, ctx: IndexingContext = None) -> _lib.Table:
        self.context = ctx
        f, n = self.host.shape
        return _lib.Table(self.host.ctypes.data, C.addressof(self._desc), n, TABLE_CACHED, f, f,
                          DeviceArray.mapped(self.host).elt, 0)

    # ---- Update-phase hook: called by update!(...) after the kernels, with the indexer of this batch
    def after_update(self, indexer, items, n_items):
        _lib.check(_lib.lib().etb_cache_admit(C.byref(indexer.view), items, n_items, self.min_count,
                                              C.c_void_p(current_stream_ptr())))

    # ---- cache management
    def cached_rows(self) -> int:
        return min(int(self._cursor.item()), self.capacity)

    def flush(self):
        """write the cached rows back to the host table (the cache stays valid)"""
        d = self.descriptor()
        _lib.check(_lib.lib().etb_cache_flush(C.byref(d), C.c_void_p(current_stream_ptr())))
        return self

    def clear(self):
        """flush, then forget every cached row"""
        self.flush()
        self._slot_of_row.fill_(-1)
        self._cursor.zero_()
        return self

    def hit_rate(self, I) -> float:
        """fraction of the occurrences in I (1-based indices, any shape) whose row is cached now"""
        idx = torch.as_tensor(np.asarray(I).reshape(-1) - 1, device="cuda", dtype=torch.int64)
        return float((self._slot_of_row[idx] >= 0).float().mean().item()) if idx.numel() else 0.0

    def to_numpy(self):
        self.flush()
        torch.cuda.current_stream().synchronize()
        return np.asfortranarray(self.host.copy())

    def _scalar(self, col, row, v=None):
        self.flush()
        torch.cuda.current_stream().synchronize()
        if v is None:
            return self.host[row - 1, col - 1].item()
        raise NotImplementedError("scalar stores into a CachedEmbedding: assign through the host array before caching")

    def __repr__(self):
        f, n = self.size()
        return f"{f}x{n} CachedEmbedding{{{self.lookup_type}, {self.dtype}}} ({self.capacity} HBM slots)"
