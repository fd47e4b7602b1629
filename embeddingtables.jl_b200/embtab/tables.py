"""Table types of the host mirror: the reference's L0 storage layer on HBM.

Mirrors (names, argument meaning, error behaviour):
  AbstractEmbeddingTable / Static / Dynamic / featuresize / IndexingContext / columnpointer /
  example                                   reference src/EmbeddingTables.jl:49-118
  SimpleEmbedding                           reference src/simple.jl:2-56
  SplitEmbedding                            reference src/split.jl:3-86
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .darray import DeviceArray, as_device


class ArgumentError(ValueError):
    """Julia's ArgumentError (thrown by the reference's constructors, src/simple.jl:10-24)."""


class AbstractLookupType:
    pass


class Dynamic(AbstractLookupType):
    """Feature size not known statically (reference src/EmbeddingTables.jl:61)."""

    def __eq__(self, other):
        return isinstance(other, Dynamic)

    def __hash__(self):
        return hash("Dynamic")

    def __repr__(self):
        return "Dynamic"


class Static(AbstractLookupType):
    """Static{N} (reference src/EmbeddingTables.jl:62-63).  N is validated by the table ctor."""

    def __init__(self, N):
        self.N = N

    def __eq__(self, other):
        return isinstance(other, Static) and other.N == self.N

    def __hash__(self):
        return hash(("Static", self.N))

    def __repr__(self):
        return f"Static{{{self.N}}}"


class IndexingContext:
    pass


class NoContext(IndexingContext):
    pass


class Forward(IndexingContext):
    pass


class Update(IndexingContext):
    pass


def device_descriptor(table, ctx: IndexingContext):
    """table.descriptor(ctx), tolerating table types whose descriptor() takes no context"""
    fn = table.descriptor
    try:
        takes_ctx = fn.__func__.__code__.co_argcount >= 2
    except AttributeError:
        takes_ctx = True
    return fn(ctx) if takes_ctx else fn()


class AbstractEmbeddingTable:
    """AbstractEmbeddingTable{S,T}.  A subtype provides: `lookup_type` (S), `dtype` (T), `size()`,
    `columnpointer(i)`, `example()` and `descriptor()` -- the device form of `columnpointer` that
    the kernels use (README.md:288-307 extension contract, here with one extra method because a
    GPU kernel cannot call back into the host for each row address)."""

    lookup_type: AbstractLookupType
    dtype: np.dtype

    def size(self, d=None):
        raise NotImplementedError

    def columnpointer(self, i: int, ctx: IndexingContext = None) -> int:
        raise ArgumentError(f"Please explicitly define `columnpointer` for {type(self).__name__}")

    def example(self) -> DeviceArray:
        raise NotImplementedError

    def descriptor(self, ctx: IndexingContext = None) -> _lib.Table:
        """device form of columnpointer(table, i, ctx): `ctx` is Forward() from the lookups and Update() from update!
        (reference src/lookup.jl:57-161, src/sparseupdate.jl:25-122); most tables ignore it"""
        raise NotImplementedError

    # --- AbstractArray interface (reference src/EmbeddingTables.jl:144-156): scalar access via
    # columnpointer; one 1-element copy per access -- slow, tests only.
    def __getitem__(self, ij):
        i, j = ij
        f, n = self.size()
        if not (1 <= i <= f and 1 <= j <= n):
            raise IndexError(f"BoundsError: attempt to access {f}x{n} table at index [{i}, {j}]")
        return self._scalar(j, i)

    def __setitem__(self, ij, v):
        i, j = ij
        f, n = self.size()
        if not (1 <= i <= f and 1 <= j <= n):
            raise IndexError(f"BoundsError: attempt to access {f}x{n} table at index [{i}, {j}]")
        self._scalar(j, i, v)

    def __len__(self):
        f, n = self.size()
        return f * n

    def __eq__(self, other):
        if isinstance(other, AbstractEmbeddingTable):
            other = other.to_numpy()
        if isinstance(other, DeviceArray):
            other = other.numpy()
        other = np.asarray(other)
        return other.shape == self.size() and bool(np.array_equal(self.to_numpy(), other))

    __hash__ = None


def featuresize(A):
    """featuresize (reference src/EmbeddingTables.jl:71-72)."""
    if isinstance(A, AbstractEmbeddingTable):
        return A.size()[0]
    return A.shape[0]


def example(x):
    """example(table) / example(Vector{table}) (reference src/EmbeddingTables.jl:118)."""
    if isinstance(x, (list, tuple)):
        return x[0].example()
    return x.example()


def columnpointer(A, i: int, ctx: IndexingContext = None) -> int:
    """columnpointer(A, i[, ctx]) -> device address of embedding row i (1-based).
    Plain matrices: pointer + stride*(i-1) (reference src/EmbeddingTables.jl:83-90)."""
    if isinstance(A, DeviceArray):
        return A.ptr + A.ld * A.itemsize * (i - 1)
    return A.columnpointer(i, ctx)


def _check_static(S, nrows_in_matrix):
    if isinstance(S, Static):
        if not isinstance(S.N, (int, np.integer)) or isinstance(S.N, bool):
            raise ArgumentError(
                f"Expected the type parameter for `Static{{N}}` to be an Int. Instead, it's a {type(S.N).__name__}!")
        if S.N != nrows_in_matrix:
            raise ArgumentError(
                "Parameter `N` should match the number of rows in the passed Matrix. "
                f"Instead, `N = {S.N}` while `size(A,1) = {nrows_in_matrix}`.")


class SimpleEmbedding(AbstractEmbeddingTable):
    """SimpleEmbedding{S}(A): thin wrapper over an HBM-resident column-major matrix.

    SimpleEmbedding(A)              -> Dynamic          (reference src/simple.jl:7-8)
    SimpleEmbedding(A, Static(N))   -> Static{N}, ArgumentError if N is not an Int or N != size(A,1)
                                       (reference src/simple.jl:9-27)
    `A` may be a host array in Julia shape (featuresize, nrows) -- uploaded -- or a DeviceArray.
    """

    def __init__(self, A, lookup_type: AbstractLookupType = None):
        S = lookup_type if lookup_type is not None else Dynamic()
        shape = A.shape
        if len(shape) != 2:
            raise ArgumentError("SimpleEmbedding wraps a matrix")
        _check_static(S, shape[0])
        self.data = as_device(A)
        self.lookup_type = S
        self.dtype = self.data.dtype

    def size(self, d=None):
        return self.data.size(d)

    def parent(self):
        return self.data

    def pointer(self):
        return self.data.ptr

    def columnpointer(self, i, ctx=None):
        # Static: pointer + (i-1)*N*sizeof(T) (src/simple.jl:53-55); Dynamic honours the stride (:52)
        if isinstance(self.lookup_type, Static):
            return self.data.ptr + (i - 1) * self.lookup_type.N * self.data.itemsize
        return self.data.ptr + (i - 1) * self.data.ld * self.data.itemsize

    def example(self):
        return self.data

    def descriptor(self, ctx=None):
        d = self.data
        ld = self.lookup_type.N if isinstance(self.lookup_type, Static) else d.ld
        return _lib.Table(d.ptr, None, d.shape[1], 0, d.shape[0], ld, d.elt, 0)

    def zeros(self):
        """Base.zeros(x::SimpleEmbedding) (reference src/simple.jl:30-34)."""
        return SimpleEmbedding(DeviceArray.zeros(self.data.shape, self.dtype), self.lookup_type)

    def to_numpy(self):
        return self.data.numpy()

    def assign(self, A):
        """`table .= A`."""
        self.data.copy_from(A)
        return self

    def _scalar(self, col, row, v=None):
        k = self.data.offset + (col - 1) * self.data.ld + (row - 1)
        if v is None:
            return self.data.buf[k].item()
        self.data.buf[k] = v

    def __repr__(self):
        f, n = self.size()
        return f"{f}x{n} SimpleEmbedding{{{self.lookup_type}, {self.dtype}}}"


class SplitEmbedding(AbstractEmbeddingTable):
    """SplitEmbedding(A, cols_per_shard=1): the table sharded into equal-width chunk matrices, the
    last one ragged (reference src/split.jl:11-26; always Static{size(A,1)}, :25).
    SplitEmbedding.undef(S, T, featuresize, ncols, cols_per_shard) is the `undef` constructor
    (:29-46).  The device form is an HBM array of chunk base pointers."""

    def __init__(self, A, cols_per_shard: int = 1, _chunks=None, _lookup_type=None, _dtype=None, _fs=None):
        if _chunks is None:
            host = A.numpy() if isinstance(A, DeviceArray) else np.asarray(A)
            if host.ndim != 2:
                raise ArgumentError("SplitEmbedding wraps a matrix")
            ncols = host.shape[1]
            self.data = [DeviceArray.from_numpy(host[:, s:min(s + cols_per_shard, ncols)])
                         for s in range(0, ncols, cols_per_shard)]
            self.lookup_type = Static(host.shape[0])
            self.dtype = host.dtype
            fs = host.shape[0]
        else:
            self.data, self.lookup_type, self.dtype, fs = _chunks, _lookup_type, np.dtype(_dtype), _fs
        self.matrixsize = (fs, int(cols_per_shard))
        ptrs = np.array([c.ptr for c in self.data], dtype=np.uint64)
        self._chunk_ptrs = torch.from_numpy(ptrs.view(np.int64)).to("cuda")

    @staticmethod
    def undef(S, T, featuresize: int, ncols: int, cols_per_shard: int = 1):
        if isinstance(S, Static):
            assert S.N == featuresize  # __compare, reference src/split.jl:49
        chunks = [DeviceArray.empty((featuresize, min(s + cols_per_shard, ncols) - s), T)
                  for s in range(0, ncols, cols_per_shard)]
        return SplitEmbedding(None, cols_per_shard, _chunks=chunks, _lookup_type=S, _dtype=T, _fs=featuresize)

    def size(self, d=None):
        nrows = self.matrixsize[0]
        ncols = self.matrixsize[1] * (len(self.data) - 1) + self.data[-1].shape[1]  # src/split.jl:70-74
        s = (nrows, ncols)
        return s if d is None else s[d - 1]

    def columnpointer(self, i, ctx=None):
        chunk, col = divmod(i - 1, self.matrixsize[1])  # _divrem_index, src/split.jl:59-65
        c = self.data[chunk]
        return c.ptr + col * c.ld * c.itemsize

    def example(self):
        return self.data[0]

    def descriptor(self, ctx=None):
        fs, shard = self.matrixsize
        return _lib.Table(None, self._chunk_ptrs.data_ptr(), self.size()[1], shard, fs, fs,
                          self.data[0].elt, 0)

    def to_numpy(self):
        return np.asfortranarray(np.concatenate([c.numpy() for c in self.data], axis=1))

    def assign(self, A):
        A = np.asarray(A, dtype=self.dtype)
        shard = self.matrixsize[1]
        for k, c in enumerate(self.data):
            c.copy_from(A[:, k * shard:k * shard + c.shape[1]])
        return self

    def _scalar(self, col, row, v=None):
        chunk, within = divmod(col - 1, self.matrixsize[1])
        c = self.data[chunk]
        k = c.offset + within * c.ld + (row - 1)
        if v is None:
            return c.buf[k].item()
        c.buf[k] = v

    def __repr__(self):
        f, n = self.size()
        return f"{f}x{n} SplitEmbedding{{{self.lookup_type}, {self.dtype}}} in {len(self.data)} chunks"


def zeros(x: SimpleEmbedding):
    return x.zeros()
