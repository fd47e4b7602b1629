"""ctypes front-end of the CPU oracle (oracle/embtab_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product package
(embeddingtables.jl_b200/) never does.

Conventions are Julia's, so the parity tests read like the reference's own tests:
  * a table is a Fortran-ordered numpy matrix of shape (featuresize, nrows);
  * indices are 1-based int64; a vector is a non-reducing lookup, a (bag, batch) matrix a
    pooled one (reference README.md:13-25);
  * outputs are Fortran-ordered (featuresize, batch).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libembtab_oracle.so")

F32, F64, I32, I64 = 0, 1, 2, 3
_ELT = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32,
        np.dtype(np.int64): I64}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "embtab_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libembtab_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Table(C.Structure):
    _fields_ = [("base", C.c_void_p), ("chunks", C.POINTER(C.c_void_p)), ("nrows", C.c_int64),
                ("shard_rows", C.c_int64), ("dim", C.c_int32), ("ld", C.c_int32),
                ("elt", C.c_int32), ("is_static", C.c_int32)]


class _LookupItem(C.Structure):
    _fields_ = [("table", _Table), ("idx", C.c_void_p), ("dst", C.c_void_p),
                ("ld_dst", C.c_int64), ("batch", C.c_int64), ("bag", C.c_int64),
                ("ld_idx", C.c_int64)]


class _UpdateItem(C.Structure):
    _fields_ = [("table", _Table), ("delta", C.c_void_p), ("ld_delta", C.c_int64),
                ("idx", C.c_void_p), ("batch", C.c_int64), ("bag", C.c_int64),
                ("ld_idx", C.c_int64), ("cum_col", C.c_void_p), ("cum_off", C.c_void_p),
                ("map", C.c_void_p), ("nnz", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.etbo_histogram_dense.restype = C.c_int64
        _lib.etbo_index_dense.restype = C.c_int64
        _lib.etbo_index_sparse.restype = C.c_int64
        _lib.etbo_uses_avx512.restype = C.c_int
        _lib.etbo_allowed_cpus.restype = C.c_int
    return _lib


def _fmat(a, dtype=None):
    a = np.asarray(a, dtype=dtype)
    if a.ndim <= 1:
        return np.ascontiguousarray(a)
    if a.flags.f_contiguous:
        return a
    if a.ndim == 2 and a.strides[0] == a.itemsize and a.strides[1] >= a.shape[0] * a.itemsize:
        return a  # a row-slice view of a Fortran matrix: columns contiguous, ld > rows
    return np.asfortranarray(a)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _ld(a):
    """leading dimension (elements) of a Fortran-ordered 2-D array or view."""
    if a.ndim == 1:
        return a.shape[0]
    assert a.shape[0] == 1 or a.strides[0] == a.itemsize, "feature vectors must be contiguous"
    return max(a.strides[1] // a.itemsize, a.shape[0]) if a.shape[1] > 1 else a.shape[0]


class Table:
    """Oracle-side SimpleEmbedding / SplitEmbedding (reference src/simple.jl, src/split.jl)."""

    def __init__(self, data, static=False, cols_per_shard=None):
        data = _fmat(data)
        self.dim, self.nrows = data.shape
        self.dtype = data.dtype
        self.static = bool(static)
        self.cols_per_shard = cols_per_shard
        if cols_per_shard is None:
            self.data = data
            self.chunks = None
        else:  # SplitEmbedding(A, cols_per_shard): reference src/split.jl:11-26, always Static
            self.static = True
            self.chunks = [np.asfortranarray(data[:, s:s + cols_per_shard].copy())
                           for s in range(0, self.nrows, cols_per_shard)]
            self.data = None
        self._keep = None

    def c(self) -> _Table:
        t = _Table()
        t.nrows, t.dim, t.elt, t.is_static = self.nrows, self.dim, _ELT[self.dtype], int(self.static)
        if self.chunks is None:
            t.base, t.chunks, t.shard_rows, t.ld = self.data.ctypes.data, None, 0, _ld(self.data)
        else:
            arr = (C.c_void_p * len(self.chunks))(*[c.ctypes.data for c in self.chunks])
            self._keep = arr
            t.base, t.chunks, t.shard_rows, t.ld = None, arr, self.cols_per_shard, self.dim
        return t

    def dense(self) -> np.ndarray:
        return self.data if self.chunks is None else np.asfortranarray(np.concatenate(self.chunks, axis=1))

    def copy(self) -> "Table":
        return Table(self.dense().copy(order="F"), self.static, self.cols_per_shard)


def _idx(I):
    return _fmat(I, np.int64)


def lookup(table: Table, I, out=None) -> np.ndarray:
    """lookup(A, I): reference src/lookup.jl:35-40."""
    I = _idx(I)
    batch = I.shape[-1]
    if out is None:
        out = np.empty((table.dim, batch), dtype=table.dtype, order="F")
    t = table.c()
    if I.ndim == 1:
        lib().etbo_gather(_ptr(out), C.c_int64(_ld(out)), C.byref(t), _ptr(I), C.c_int64(batch))
    else:
        lib().etbo_pooled_sum(_ptr(out), C.c_int64(_ld(out)), C.byref(t), _ptr(I),
                              C.c_int64(I.shape[0]), C.c_int64(batch), C.c_int64(_ld(I)))
    return out


def colwrap(tables, I):
    """colwrap, reference src/lookup.jl:195-213: list -> as is; N-d array -> last-dim slices."""
    if isinstance(I, (list, tuple)):
        return [_idx(i) for i in I]
    I = np.asarray(I)
    return [_idx(I[..., t]) for t in range(I.shape[-1])]


def maplookup(strategy: str, tables, I, prependrows=0, nthreads=1, worksize_div=8, out=None):
    """maplookup(strategy, tables, I), reference src/lookup.jl:220-371.
    strategy in {"default", "simple_parallel", "preallocation"}."""
    Is = colwrap(tables, I)
    code = {"default": 0, "simple_parallel": 1, "preallocation": 2}[strategy]
    batch = Is[0].shape[-1]
    items = (_LookupItem * len(tables))()
    keep = []
    if code == 2:
        total = prependrows + sum(t.dim for t in tables)
        if out is None:
            out = np.empty((total, batch), dtype=tables[0].dtype, order="F")
        outs = out
        off = prependrows
    else:
        outs = [np.empty((t.dim, batch), dtype=t.dtype, order="F") for t in tables]
    for k, (t, i) in enumerate(zip(tables, Is)):
        it = items[k]
        it.table = t.c()
        it.idx = i.ctypes.data
        it.batch = batch
        it.bag = 0 if i.ndim == 1 else i.shape[0]
        it.ld_idx = 0 if i.ndim == 1 else _ld(i)
        if code == 2:
            it.dst = out.ctypes.data + off * out.itemsize
            it.ld_dst = _ld(out)
            off += t.dim
        else:
            it.dst = outs[k].ctypes.data
            it.ld_dst = t.dim
        keep.append(i)
    lib().etbo_maplookup(items, C.c_int(len(tables)), C.c_int(code), C.c_int(nthreads),
                         C.c_int(worksize_div))
    return outs


def histogram_dense(A, maxindex):
    """histogram!(array, A), reference src/utils.jl:131-134,154-167 -> (order[], count[])."""
    A = _idx(A).ravel(order="F")
    order = np.zeros(maxindex, np.int64)
    count = np.zeros(maxindex, np.int64)
    nnz = lib().etbo_histogram_dense(_ptr(A), C.c_int64(A.size), C.c_int64(maxindex), _ptr(order),
                                     _ptr(count))
    return nnz, order, count


def index(A, maxindex, dense=False):
    """index!(Indexer, A, maxindex), reference src/utils.jl:306-314.
    Returns (cumulative, map): cumulative = list of (col, offset) incl. terminator, 1-based."""
    A = _idx(A)
    bag = 0 if A.ndim == 1 else A.shape[0]
    flat = A.ravel(order="F")
    n = flat.size
    cum_col = np.zeros(n + 1, np.int64)
    cum_off = np.zeros(n + 1, np.int64)
    mp = np.zeros(n, np.int64)
    fn = lib().etbo_index_dense if dense else lib().etbo_index_sparse
    nnz = fn(_ptr(flat), C.c_int64(n), C.c_int64(bag), C.c_int64(maxindex), _ptr(cum_col),
             _ptr(cum_off), _ptr(mp))
    return list(zip(cum_col[:nnz + 1].tolist(), cum_off[:nnz + 1].tolist())), mp


def buckets(cumulative, mp):
    """{row: [delta columns in occurrence order]} -- the order-insensitive view of an Indexer
    (bucket order differs between the reference (first-seen) and the GPU sort (ascending))."""
    out = {}
    for (col, off), (_, nxt) in zip(cumulative[:-1], cumulative[1:]):
        out[col] = mp[off - 1:nxt - 1].tolist()
    return out


def indexer_view(cum_len, num_splits, this_split):
    s, e = C.c_int64(), C.c_int64()
    lib().etbo_indexer_view(C.c_int64(cum_len), C.c_int64(num_splits), C.c_int64(this_split),
                            C.byref(s), C.byref(e))
    return s.value, e.value


def update(table: Table, delta, I, eta, dense=False, split=None):
    """update!(Descent(eta), table, SparseEmbeddingUpdate(delta, I)), reference
    src/sparseupdate.jl:160-178.  split=(num_splits, this_split) applies one IndexerView."""
    delta = _fmat(delta)
    cumulative, mp = index(I, table.nrows, dense)
    cum_col = np.array([c for c, _ in cumulative], np.int64)
    cum_off = np.array([o for _, o in cumulative], np.int64)
    if split is None:
        e0, e1 = 0, len(cumulative) - 1
    else:
        s, e = indexer_view(len(cumulative), *split)
        e0, e1 = s - 1, e - 1
    t = table.c()
    lib().etbo_update(C.byref(t), _ptr(delta), C.c_int64(_ld(delta)), _ptr(cum_col), _ptr(cum_off),
                      _ptr(mp), C.c_int64(e0), C.c_int64(e1), C.c_double(eta))


def update_ensemble(tables, deltas, Is, eta, num_splits=4, nthreads=1, dense=False, scratch=None):
    """update!(opt, tables, grads, indexers; num_splits, nthreads), reference
    src/sparseupdate.jl:199-238.  `scratch` = preallocated Indexer storage (see alloc_indexers)."""
    n = len(tables)
    items = (_UpdateItem * n)()
    keep = []
    if scratch is None:
        scratch = alloc_indexers(Is)
    for k in range(n):
        i = _idx(Is[k])
        d = _fmat(deltas[k])
        it = items[k]
        it.table = tables[k].c()
        it.delta, it.ld_delta = d.ctypes.data, _ld(d)
        it.idx, it.batch = i.ctypes.data, i.shape[-1]
        it.bag = 0 if i.ndim == 1 else i.shape[0]
        it.ld_idx = it.bag
        cc, co, mp = scratch[k]
        it.cum_col, it.cum_off, it.map = cc.ctypes.data, co.ctypes.data, mp.ctypes.data
        keep += [i, d]
    lib().etbo_update_ensemble(items, C.c_int(n), C.c_double(eta), C.c_int(num_splits),
                               C.c_int(nthreads), C.c_int(int(dense)))
    return scratch


def set_pinning(on: bool):
    """bench.py's CPU arm: pin worker thread t to the t-th CPU this process may run on"""
    lib().etbo_set_pinning(C.c_int(int(on)))


def allowed_cpus() -> int:
    return int(lib().etbo_allowed_cpus())


def fill_uniform(a: np.ndarray, seed: int, nthreads: int):
    """threaded uniform [0, 1) Float32 fill (synthetic tables of the CPU arm; first touch by pinned threads)"""
    assert a.dtype == np.float32 and (a.flags.f_contiguous or a.flags.c_contiguous)
    lib().etbo_fill_uniform(C.c_void_p(a.ctypes.data), C.c_size_t(a.size), C.c_uint64(seed), C.c_int(nthreads))
    return a


def alloc_indexers(Is):
    out = []
    for i in Is:
        n = int(np.asarray(i).size)
        out.append((np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64), np.zeros(n, np.int64)))
    return out


def uncompress(delta, I, dstcols):
    """uncompress(SparseEmbeddingUpdate(delta, I), dstcols), reference src/sparseupdate.jl:16-32."""
    delta = _fmat(delta)
    I = _idx(I)
    dst = np.zeros((delta.shape[0], dstcols), dtype=delta.dtype, order="F")
    bag = 0 if I.ndim == 1 else I.shape[0]
    lib().etbo_uncompress(_ptr(dst), C.c_int64(_ld(dst)), C.c_int32(delta.shape[0]),
                          C.c_int32(_ELT[delta.dtype]), _ptr(delta), C.c_int64(_ld(delta)), _ptr(I),
                          C.c_int64(bag), C.c_int64(I.shape[-1]), C.c_int64(bag))
    return dst


def force_portable(on: bool):
    lib().etbo_force_portable(C.c_int(int(on)))


# ------------------------------------------------------------------ half-precision extension
# ETB_F16 / ETB_BF16 tables have NO reference counterpart: the reference's tables are
# Float32/Float64/integer.  PARITY UNPINNED for this extension -- these functions DEFINE its semantics
# (include/embtab_b200.h, ETB_F16): every element is converted to Float32 (exact), sums run in Float32 in the
# reference's order (pooled: accumulator = first row, then bag order; update: 0 + members in occurrence order),
# the update epilogue is `row - eta*acc` in Float32 with separate roundings, and the result is rounded to the
# storage type once (nearest-even).  numpy's float32 ufuncs round each operation separately, like the kernels.
def lookup_lowp(data, I):
    data = np.asarray(data)
    I = np.asarray(I, dtype=np.int64)
    if I.ndim == 1:
        return np.asfortranarray(data[:, I - 1])
    acc = data[:, I[0] - 1].astype(np.float32)
    for k in range(1, I.shape[0]):
        acc = acc + data[:, I[k] - 1].astype(np.float32)
    return np.asfortranarray(acc.astype(data.dtype))


def update_lowp(data, delta, I, eta):
    """in place on `data` (featuresize x nrows, float16 / ml_dtypes.bfloat16)"""
    I = np.asarray(I, dtype=np.int64)
    bag = I.shape[0] if I.ndim == 2 else 1
    flat = I.reshape(-1, order="F")
    eta32 = np.float32(eta)
    acc = {}
    for p, r in enumerate(flat.tolist()):
        d = np.asarray(delta[:, p // bag]).astype(np.float32)
        acc[r] = (acc[r] + d) if r in acc else (np.float32(0) + d)
    for r, a in acc.items():
        row = data[:, r - 1].astype(np.float32)
        data[:, r - 1] = (row - eta32 * a).astype(data.dtype)
    return data


# ------------------------------------------------------------------ row-wise Adagrad extension
# etb_adagrad_update has NO reference counterpart either (the reference's update! exists for
# Flux.Descent only): PARITY UNPINNED; this function DEFINES the semantics stated in include/embtab_b200.h and
# repeats the kernel's fixed summation order for the sum of squares (element d lives in vector d // NE, vector v
# on lane v % G in slot v // G; each lane adds its squares slot by slot, element by element; the G lanes are then
# combined by an XOR butterfly), every operation rounded separately.
def adagrad_layout(dim, itemsize, vb=None):
    rowbytes = dim * itemsize
    if vb is None:  # dense, aligned tables and cotangents: the widest vector that divides the row
        vb = next((v for v in (16, 8) if v >= itemsize and rowbytes % v == 0), 8 if itemsize == 8 else 4)
    nvec = rowbytes // vb
    G = min(32, 1 << max(0, (nvec - 1).bit_length()))
    return nvec, G, vb // itemsize


def adagrad_update(data, state, delta, I, eta, eps, vb=None):
    """in place on `data` (featuresize x nrows) and `state` (nrows, float32 -- float64 for float64 tables)"""
    A = np.float64 if data.dtype == np.float64 else np.float32
    assert state.dtype == A
    I = np.asarray(I, dtype=np.int64)
    bag = I.shape[0] if I.ndim == 2 else 1
    flat = I.reshape(-1, order="F")
    dim = data.shape[0]
    nvec, G, NE = adagrad_layout(dim, data.dtype.itemsize, vb)
    slots = -(-nvec // G)
    assert slots <= 4, "rows of more than 128 vectors are unsupported"
    acc = {}
    for p, r in enumerate(flat.tolist()):
        d = np.asarray(delta[:, p // bag]).astype(A)
        acc[r] = (acc[r] + d) if r in acc else (A(0) + d)
    lanes = np.arange(G)
    for r, g in acc.items():
        gp = np.zeros(slots * G * NE, A)
        gp[:dim] = g
        gp = gp.reshape(slots, G, NE)
        s = np.zeros(G, A)
        for p in range(slots):
            for e in range(NE):
                s = s + gp[p, :, e] * gp[p, :, e]
        off = G >> 1
        while off:
            s = s + s[lanes ^ off]
            off >>= 1
        h = A(state[r - 1] + A(s[0] / A(dim)))
        state[r - 1] = h
        scale = A(A(eta) / A(np.sqrt(h) + A(eps)))
        row = data[:, r - 1].astype(A)
        data[:, r - 1] = (row - scale * g).astype(data.dtype)
    return data
