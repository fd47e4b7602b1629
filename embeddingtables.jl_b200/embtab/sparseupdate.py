"""SparseEmbeddingUpdate, Indexer, update! and the pullbacks of the host mirror.

Mirrors reference src/sparseupdate.jl (SparseEmbeddingUpdate :6-13, uncompress :16-32, rrule
:35-40, update! kernels :46-154, Flux compat :160-189, ensemble update :195-238), the rrules of
src/lookup.jl:247-258, 374-389, and `Slicer`/`Indexer`/`IndexerView` of src/utils.jl:50-63,
519-577.  Flux.Descent is mirrored by `Descent` (only `eta` is used, src/sparseupdate.jl:173).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .darray import DeviceArray, as_device, as_device_indices, current_stream_ptr
from .lookup import (AbstractExecutionStrategy, DefaultStrategy, PreallocationStrategy, colwrap, lookup,
                     maplookup)
from .tables import AbstractEmbeddingTable, Static, Update, device_descriptor, featuresize


class Descent:
    """Flux.Descent(eta)."""

    def __init__(self, eta=0.1):
        self.eta = float(eta)


class Adagrad:
    """Row-wise Adagrad (GPU-only extension; the reference's update! knows Flux.Descent only).

    One state element per table row, kept here per table (keyed by the table object) and created as zeros on
    first use:  h = state[k] + mean(g.^2);  state[k] = h;  A[:, k] -= eta / (sqrt(h) + eps) * g, with g the
    bucket's summed cotangent.  State arithmetic is Float32 (Float64 for Float64 tables)."""

    def __init__(self, eta=0.1, eps=1e-8):
        self.eta, self.eps = float(eta), float(eps)
        self._state = {}

    def state(self, table) -> DeviceArray:
        st = self._state.get(id(table))
        if st is None:
            acc = np.float64 if table.dtype == np.float64 else np.float32
            st = self._state[id(table)] = (DeviceArray.zeros((table.size(2),), acc), table)   # keeps the table alive
        return st[0]


class SparseEmbeddingUpdate:
    """SparseEmbeddingUpdate{S}(delta, indices): lazy, aliasing COO-like gradient
    (reference src/sparseupdate.jl:6-13).  `delta` (featuresize x batch, possibly a row-slice
    view) and `indices` are NOT copied: keep them alive and unmodified until update!."""

    def __init__(self, lookup_type, delta, indices):
        self.lookup_type = lookup_type
        self.delta = as_device(delta)
        self.indices = as_device_indices(indices)


def _update_item(table, grad: SparseEmbeddingUpdate) -> _lib.UpdateItem:
    I, d = grad.indices, grad.delta
    if I.ndim == 1:
        bag, batch, ld_idx = 0, I.shape[0], 0
    else:
        bag, batch, ld_idx = I.shape[0], I.shape[1], I.ld
    if d.dtype != table.dtype:
        raise TypeError(f"delta eltype {d.dtype} != table eltype {table.dtype}")
    if d.shape[0] < featuresize(table) or d.shape[1] < batch:
        raise ValueError(f"delta {d.shape} too small for {featuresize(table)} x {batch}")
    return _lib.UpdateItem(device_descriptor(table, Update()), d.ptr, d.ld, I.ptr, batch, bag, ld_idx, I.elt,
                           _flags(table) & _lib.UPDATE_FMA)   # the epilogue is a per-table choice


def uncompress(x: SparseEmbeddingUpdate, dstcols=None, maxindices=None) -> DeviceArray:
    """Dense gradient of a sparse update (test helper, reference src/sparseupdate.jl:16-32).
    `maxindices` stops after that many delta columns (reference :28-29)."""
    I, d = x.indices, x.delta
    if dstcols is None:
        dstcols = int(I.numpy().max())
    batch = d.shape[1] if maxindices is None else min(d.shape[1], int(maxindices))
    dst = DeviceArray.zeros((d.shape[0], dstcols), d.dtype)
    bag, ld_idx = (0, 0) if I.ndim == 1 else (I.shape[0], I.ld)
    _lib.check(_lib.lib().etb_uncompress(dst.ptr, dst.ld, d.shape[0], d.elt, d.ptr, d.ld, I.ptr, I.elt, bag,
                                         batch, ld_idx, C.c_void_p(current_stream_ptr())))
    return dst


# ------------------------------------------------------------------------------ Indexer
class AbstractIndexer:
    pass


class Indexer(AbstractIndexer):
    """Caller-owned, reusable scratch for index! (reference src/utils.jl:288-304).  On the GPU it
    is a workspace in HBM holding the sorted (row, delta column) pairs and the bucket offsets
    (include/embtab_b200.h, etb_index_view).  It grows on demand and is then reused."""

    def __init__(self):
        self.workspace = None
        self.view = None     # _lib.IndexView of the last index!
        self._items = None
        self._event = None   # set by prefetch_index: the side stream's completion event
        self._prefetched = None

    def _ensure(self, items_arr, n):
        need = C.c_size_t()
        _lib.check(_lib.lib().etb_index_workspace_bytes(items_arr, n, C.byref(need)))
        if self.workspace is None or self.workspace.numel() < need.value:
            if self.workspace is not None:
                torch.cuda.synchronize()   # kernels on other streams may still read the old workspace
            self.workspace = torch.empty(need.value, dtype=torch.uint8, device="cuda")

    # --- host-side inspection (tests): the reference's `cumulative` / `map`, order-insensitive
    def buckets(self):
        """{(slot, row): [delta columns (1-based) in occurrence order]}"""
        v = self.view
        torch.cuda.current_stream().synchronize()
        nnz = int(_peek(v.nnz, 1, np.int64)[0])
        n = v.n_total
        mp = _peek(v.map, n, np.int32)
        rec = _peek(v.records, nnz, np.dtype([("start", np.uint32), ("m0", np.int32), ("key", np.uint64)]))
        offs = np.append(rec["start"].astype(np.int64), n)
        mask = np.uint64((1 << v.row_bits) - 1)
        out = {}
        for s in range(nnz):
            k = rec["key"][s]
            members = (mp[offs[s]:offs[s + 1]] + 1).tolist()
            assert members[0] == rec["m0"][s] + 1
            out[(int(k >> np.uint64(v.row_bits)), int(k & mask) + 1)] = members
        return out


SparseIndexer = Indexer  # reference src/utils.jl:295-296: histogram flavours of the CPU algorithm;
DenseIndexer = Indexer   # the GPU sort has one flavour


class IndexerView(AbstractIndexer):
    """IndexerView(I, num_splits, this_split) (reference src/utils.jl:320-333): a sub-range of the
    buckets, for partitioned updates."""

    def __init__(self, I: Indexer, num_splits: int, this_split: int):
        self.I, self.num_splits, self.this_split = I, int(num_splits), int(this_split)


def _peek(ptr, n, dtype):
    out = np.empty(n, dtype)
    if n:
        _lib.check(_lib.lib().etb_memcpy_d2h(out.ctypes.data, ptr, out.nbytes, C.c_void_p(current_stream_ptr())))
        torch.cuda.current_stream().synchronize()
    return out


class _IndicesOnly:
    """what index! needs of a SparseEmbeddingUpdate when the cotangent does not exist yet"""

    def __init__(self, table, indices):
        self.indices = as_device_indices(indices)


def _index_item(table, g) -> _lib.UpdateItem:
    if isinstance(g, _IndicesOnly):
        I = g.indices
        bag, batch, ld_idx = (0, I.shape[0], 0) if I.ndim == 1 else (I.shape[0], I.shape[1], I.ld)
        return _lib.UpdateItem(device_descriptor(table, Update()), None, featuresize(table), I.ptr, batch, bag, ld_idx, I.elt, 0)
    return _update_item(table, g)


_SIDE_STREAM = {}


def prefetch_index(indexer: Indexer, tables, indices):
    """GPU-only extension: start index! for (tables, indices) NOW, on a side stream.

    index! depends only on the indices, not on the cotangent, so it can overlap the forward pass (and,
    for host-resident data, the PCIe copies).  A following update!(opt, tables, grads, [indexer]) on the
    same tables and indices waits for the side stream and skips its own index! phase."""
    single = isinstance(tables, AbstractEmbeddingTable)
    tables = [tables] if single else list(tables)
    Is = [indices] if single else list(colwrap(indices))
    dev = torch.cuda.current_device()
    side = _SIDE_STREAM.get(dev)
    if side is None:
        # normal priority (measured): with a high-priority side stream the sort's scatter CTAs (4 x 64
        # registers x 256 threads fill an SM's register file) crowd the forward's CTAs out: 3.56 vs 3.41 ms
        side = _SIDE_STREAM[dev] = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        index_(indexer, tables, [_IndicesOnly(t, i) for t, i in zip(tables, Is)])
        indexer._event = side.record_event()
    indexer._prefetched = [i.ptr for i in (g.indices for g in indexer._grads_keepalive)]
    return indexer


def index_(indexer: Indexer, tables, grads):
    """index!(indexer, indices, maxindex) for one table or an ensemble sharing one Indexer."""
    if isinstance(tables, AbstractEmbeddingTable):
        tables, grads = [tables], [grads]
    items = [_index_item(t, g) for t, g in zip(tables, grads)]
    indexer._grads_keepalive = list(grads)
    indexer._event, indexer._prefetched = None, None
    arr = (_lib.UpdateItem * len(items))(*items)
    indexer._ensure(arr, len(items))
    view = _lib.IndexView()
    _lib.check(_lib.lib().etb_index(indexer.workspace.data_ptr(), indexer.workspace.numel(), arr, len(items),
                                    C.byref(view), C.c_void_p(current_stream_ptr())))
    indexer.view, indexer._items = view, arr
    return indexer


# ------------------------------------------------------------------------------ update!
# Reduction order of update!.  "strict" (the default): every bucket is accumulated strictly in occurrence order
# like the reference (bit-identical to it); a bucket of more than 128 members is streamed through shared memory by
# one CTA (long_strict_kernel).  "split": such buckets (hot Zipf rows) are summed as 128-member chunks combined in
# chunk order -- deterministic and atomics-free, but a different association (within the north star's 1e-5
# tolerance); buckets of up to 128 members are identical in both modes.
DEFAULT_ORDER = "strict"
_ORDER = {"mode": DEFAULT_ORDER}


def set_update_order(mode: str):
    """'strict' (default: bit-identical to the reference's sequential accumulation) or 'split'."""
    if mode not in ("split", "strict"):
        raise ValueError(mode)
    _ORDER["mode"] = mode


def _flags(table) -> int:
    """The reference's @generated dispatch (src/sparseupdate.jl:131-154, src/simd.jl:5-12): the
    specialised FMA kernel for Static{N} Float32 tables with N*4 <= 512 and N % 16 == 0, the
    generic two-rounding kernel otherwise."""
    S = table.lookup_type
    f = _lib.UPDATE_SPLIT_LONG if _ORDER["mode"] == "split" else 0
    if (isinstance(S, Static) and table.dtype == np.float32 and S.N * 4 <= 512 and S.N % 16 == 0):
        f |= _lib.UPDATE_FMA
    return f


def _apply(tables, grads, indexer, eta, opt=None):
    """update!(table, update, indexer, alpha): apply an already-indexed update.  The cotangent applied is
    the one of the updates passed HERE (reference src/sparseupdate.jl:131-154 reads `update.delta`); the
    indexer only has to hold index! of the same index arrays."""
    items = [_update_item(t, g) for t, g in zip(tables, grads)]
    if isinstance(indexer, IndexerView):
        base = indexer.I
        view = _lib.IndexView.from_buffer_copy(base.view)
        view.num_splits, view.this_split = indexer.num_splits, indexer.this_split
    else:
        base, view = indexer, indexer.view
    flags = _lib.UPDATE_SPLIT_LONG if _ORDER["mode"] == "split" else 0   # FMA travels per item
    stream = C.c_void_p(current_stream_ptr())
    base._items = (_lib.UpdateItem * len(items))(*items)
    if isinstance(opt, Adagrad):
        states = (C.c_void_p * len(items))(*[opt.state(t).ptr for t in tables])
        _lib.check(_lib.lib().etb_adagrad_update(C.byref(view), base._items, states, len(items), opt.eta, opt.eps,
                                                 flags, stream))
    else:
        _lib.check(_lib.lib().etb_sgd_update(C.byref(view), base._items, len(items), float(eta), flags, stream))
    # Update-phase hook of table types that react to the access phase (host-tier tables admit the batch's hot rows
    # into their HBM cache from the bucket records index! just produced)
    hooks = [t for t in tables if hasattr(t, "after_update")]
    if hooks:
        hooks[0].after_update(base, base._items, len(items))


def update_table_(table, update: SparseEmbeddingUpdate, indexer, alpha, nontemporal=True, *args):
    """update!(table, update, indexer::AbstractIndexer, alpha) (src/sparseupdate.jl:46-55,131-154):
    `indexer` must already hold index!(indexer, update.indices, ...)."""
    _apply([table], [update], indexer, alpha)


def update_(opt, table, grad, indexer=None, nontemporal=True, *args, num_splits=4, nthreads=None,
            scratchspaces=None, telemetry_cb=None):
    """update!(opt::Descent, table, grad, [indexer], [Val(nontemporal)]) for one table
    (src/sparseupdate.jl:160-178) and update!(opt, tables, grads, indexers; num_splits, nthreads,
    scratchspaces, telemetry_cb) for an ensemble (:199-238).  Returns None.
    `nontemporal`, `num_splits`, `nthreads`, `scratchspaces` are CPU tuning knobs: accepted,
    ignored.  `telemetry_cb` is called between the index and the update phases like the reference."""
    if isinstance(table, AbstractEmbeddingTable):
        if indexer is None:
            indexer = Indexer()
        if not _consume_prefetch(indexer, [table], [grad]):
            index_(indexer, table, grad)
        _apply([table], [grad], indexer, opt.eta, opt)  # convert(eltype(table), opt.eta) happens in the kernel
        return None
    tables, grads = list(table), list(grad)
    indexers = indexer if indexer is not None else [Indexer()]
    ix = indexers[0] if isinstance(indexers, (list, tuple)) else indexers
    if not _consume_prefetch(ix, tables, grads):
        index_(ix, tables, grads)  # one batched sort for every table (the @batch index! phase, :211-213)
    if telemetry_cb is not None:
        telemetry_cb()
    _apply(tables, grads, ix, opt.eta, opt)
    return None


def _consume_prefetch(indexer: Indexer, tables, grads) -> bool:
    """True if `indexer` holds a prefetch_index result for exactly these index arrays: wait for it."""
    if indexer._prefetched is None or indexer._prefetched != [g.indices.ptr for g in grads]:
        return False
    torch.cuda.current_stream().wait_event(indexer._event)
    indexer._event, indexer._prefetched = None, None   # _apply builds the items from the real cotangents
    return True


def ensemble_update(nthreads: int):
    return [Indexer() for _ in range(nthreads)]


# ------------------------------------------------------------------------------ pullbacks
class Slicer:
    """Slicer(current_index, concat_dim, array): successive row-slice views of the concatenated
    cotangent (reference src/utils.jl:50-63).  The reference's call operator increments a local
    copy of `current_index`, so it never advances; its own test (test/map.jl:153-177) requires
    per-table slices, which is what this implements.  concat_dim is 1 (rows)."""

    def __init__(self, current_index: int, concat_dim: int, captured_array: DeviceArray):
        assert concat_dim == 1
        self.current_index, self.concat_dim, self.captured_array = int(current_index), 1, captured_array

    def __call__(self, sz: int) -> DeviceArray:
        lo = self.current_index - 1
        self.current_index += int(sz)
        return self.captured_array.rows(lo, lo + int(sz))


def rrule(f, *args, **kw):
    """ChainRulesCore.rrule for lookup and maplookup.  Returns (value, pullback); pullbacks are
    lazy -- they only package (cotangent, indices) into SparseEmbeddingUpdates.

    rrule(lookup, A, I)                          reference src/sparseupdate.jl:35-40
    rrule(maplookup, strategy, A, I)             reference src/lookup.jl:247-258
    rrule(maplookup, PreallocationStrategy, A, I) reference src/lookup.jl:374-389
    """
    if f is lookup:
        A, I = args
        I = as_device_indices(I)
        S = A.lookup_type

        def lookup_pullback(delta):
            return (None, SparseEmbeddingUpdate(S, delta, I), None)

        return lookup(A, I), lookup_pullback
    if f is maplookup:
        if isinstance(args[0], AbstractExecutionStrategy):
            strategy, A, I = args
        else:
            strategy = DefaultStrategy()
            A, I = args
        A = list(A)
        S = A[0].lookup_type
        Is = list(colwrap(I))
        result = maplookup(strategy, A, Is, **kw)
        if isinstance(strategy, PreallocationStrategy):
            def maplookup_pullback(delta):
                fslice = Slicer(strategy.prependrows + 1, 1, as_device(delta))
                ds = [SparseEmbeddingUpdate(S, fslice(featuresize(y)), x) for y, x in zip(A, Is)]
                return (None, None, ds, None)
        else:
            def maplookup_pullback(deltas):
                return (None, None, [SparseEmbeddingUpdate(S, d, i) for d, i in zip(deltas, Is)], None)
        return result, maplookup_pullback
    raise TypeError(f"no rrule for {f}")


def pullback(f, *args, **kw):
    """Zygote._pullback(f, args...) restricted to this path: (y, back) with back(delta) ->
    (None, gradients...) like the reference's tests use it (test/update.jl:20-43)."""
    return rrule(f, *args, **kw)
