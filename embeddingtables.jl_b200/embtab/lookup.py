"""lookup / lookup! / maplookup / maplookup! of the host mirror.

Mirrors reference src/lookup.jl: `destination` (:19-22), `lookup`/`lookup!` (:35-43, dispatch
:90-102, :167-182), `ColumnWrap`/`colwrap` (:195-213), the three execution strategies
(:220-241, :262-276, :284-371).  Julia's `f!` is spelled `f_` here.

Every strategy lowers to ONE etb_maplookup call (one kernel launch when the tables share dim and
dtype): on a GPU the three strategies differ only in where the destinations point.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .darray import DeviceArray, as_device_indices, current_stream_ptr
from .tables import AbstractEmbeddingTable, Forward, device_descriptor, example, featuresize


# ------------------------------------------------------------------------------ destinations
def _trailing_size(I) -> int:
    return I.shape[-1]


def destination(A: AbstractEmbeddingTable, I) -> DeviceArray:
    """similar(example(A), eltype(A), featuresize(A), size(I)[end]) (src/lookup.jl:19-22)."""
    return example(A).similar(A.dtype, (featuresize(A), _trailing_size(I)))


def _item(table, I: DeviceArray, dst: DeviceArray) -> _lib.LookupItem:
    if I.ndim == 1:
        bag, batch, ld_idx = 0, I.shape[0], 0
    elif I.ndim == 2:
        bag, batch, ld_idx = I.shape[0], I.shape[1], I.ld
    else:
        raise TypeError("indices must be a vector (gather) or a matrix (pooled sum)")
    if dst.ndim != 2 or dst.shape[0] < featuresize(table) or dst.shape[1] < batch:
        raise ValueError(f"destination {dst.shape} too small for {featuresize(table)} x {batch}")
    if dst.dtype != table.dtype:
        raise TypeError(f"destination eltype {dst.dtype} != table eltype {table.dtype}")
    return _lib.LookupItem(device_descriptor(table, Forward()), I.ptr, dst.ptr, dst.ld, batch, bag, ld_idx, I.elt, 0)


def _run(items):
    arr = (_lib.LookupItem * len(items))(*items)
    _lib.check(_lib.lib().etb_maplookup(arr, len(items), C.c_void_p(current_stream_ptr())))


# ------------------------------------------------------------------------------ single table
def lookup_(dst: DeviceArray, src: AbstractEmbeddingTable, indices) -> DeviceArray:
    """lookup!(dst, src, indices): gather for a vector, ordered pooled sum for a matrix."""
    I = as_device_indices(indices)
    if _trailing_size(I) > 0:
        _run([_item(src, I, dst)])
    return dst


def lookup(A: AbstractEmbeddingTable, I) -> DeviceArray:
    """lookup(A, I) = lookup!(destination(A, I), A, I) (src/lookup.jl:35-40)."""
    I = as_device_indices(I)
    return lookup_(destination(A, I), A, I)


# ------------------------------------------------------------------------------ ColumnWrap
class ColumnWrap:
    """Treat an N-d index array as a vector of its last-dimension slices (src/lookup.jl:195-208)."""

    def __init__(self, array: DeviceArray):
        self.array = array

    def __len__(self):
        return self.array.shape[-1]

    def __getitem__(self, i):  # 0-based here
        return self.array.lastdim(i)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def colwrap(x):
    """colwrap (src/lookup.jl:211-213): a vector of index arrays passes through, an N-d array is
    wrapped."""
    if isinstance(x, ColumnWrap):
        return x
    if isinstance(x, (list, tuple)):
        return [as_device_indices(i) for i in x]
    return ColumnWrap(as_device_indices(x))


def _batchsize(x) -> int:
    """_batchsize (src/lookup.jl:296-299)."""
    if isinstance(x, ColumnWrap):
        x = x.array
    if isinstance(x, (list, tuple)):
        return _trailing_size(x[0])
    if x.ndim == 1:
        return 1
    return x.shape[-2]


# ------------------------------------------------------------------------------ strategies
class AbstractExecutionStrategy:
    pass


class DefaultStrategy(AbstractExecutionStrategy):
    """reference: serial map(lookup!, ...) (src/lookup.jl:220-241)."""


class SimpleParallelStrategy(AbstractExecutionStrategy):
    """reference: static table-parallel threads (src/lookup.jl:262-276)."""


class PreallocationStrategy(AbstractExecutionStrategy):
    """PreallocationStrategy{T}(prependrows): fuse the lookups with the concatenation
    (src/lookup.jl:284-291).  `eltype` = the {T} override of the output element type
    (`_select_eltype`, :293-294); when it differs from the tables' eltype the lookups are
    converted into the output matrix after the kernel (see maplookup_)."""

    def __init__(self, prependrows: int = 0, eltype=None):
        self.prependrows = int(prependrows)
        self.eltype = None if eltype is None else np.dtype(eltype)


def maplookup(*args, **kw):
    """maplookup([strategy], tables, I) (src/lookup.jl:221-231, 305-314)."""
    if isinstance(args[0], AbstractExecutionStrategy):
        strategy, x, I0 = args
    else:
        strategy = DefaultStrategy()
        x, I0 = args
    x = list(x)
    I = colwrap(I0)
    if isinstance(strategy, PreallocationStrategy):
        T = x[0].dtype if strategy.eltype is None else strategy.eltype
        nrows = strategy.prependrows + sum(featuresize(t) for t in x)
        dst = example(x).similar(T, (nrows, _batchsize(I)))  # first `prependrows` rows uninitialised
        return maplookup_(strategy, dst, x, I, **kw)
    y = [destination(t, i) for t, i in zip(x, I)]
    return maplookup_(strategy, y, x, I)


def maplookup_(strategy: AbstractExecutionStrategy, out, x, I0, worksize_div: int = 8):
    """maplookup!(strategy, out, tables, I).  Default/SimpleParallel: `out` is a vector of
    matrices; Preallocation: one (prependrows + sum(featuresize)) x batch matrix whose first
    `prependrows` rows are left untouched (src/lookup.jl:316-371).  `worksize_div` is a CPU
    load-balancing knob: accepted and ignored."""
    x = list(x)
    I = colwrap(I0)
    if len(I) != len(x):
        raise ValueError(f"{len(x)} tables but {len(I)} index arrays")
    items = []
    if isinstance(strategy, PreallocationStrategy) and x and out.dtype != x[0].dtype:
        # PreallocationStrategy{U} with U != eltype(tables) (`_select_eltype`, src/lookup.jl:293-294, 312): the kernels
        # write the tables' element type, so the lookups land in a scratch matrix of that type and are converted into
        # `out` below its prepend rows.  Exact for non-reducing lookups (convert(U, x)); a pooled sum is formed in the
        # tables' type and converted once, where the reference's generic path would accumulate in promote_type(U, T).
        rows = sum(featuresize(t) for t in x)
        batch = _batchsize(I)
        tmp = example(x).similar(x[0].dtype, (rows, batch))
        maplookup_(PreallocationStrategy(0), tmp, x, I)
        view = torch.as_strided(out.buf, (batch, rows), (out.ld, 1), out.offset + strategy.prependrows)
        view.copy_(tmp.buf[tmp.offset:tmp.offset + rows * batch].view(batch, rows))
        return out
    if isinstance(strategy, PreallocationStrategy):
        off = strategy.prependrows
        batch = _batchsize(I)
        for t, i in zip(x, I):
            f = featuresize(t)
            if _trailing_size(i) > 0:
                items.append(_item(t, i, out.rows(off, off + f).cols(0, batch)))
            off += f
    else:
        for o, t, i in zip(out, x, I):
            if _trailing_size(i) > 0:
                items.append(_item(t, i, o))
    if items:
        _run(items)
    return out
