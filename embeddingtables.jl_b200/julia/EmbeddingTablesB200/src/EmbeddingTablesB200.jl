# EmbeddingTablesB200 -- Julia host layer over libembtab_b200.so (include/embtab_b200.h).
#
# EXPERIMENTAL AND NOT EXECUTED IN THIS REPOSITORY'S ENVIRONMENT: the build image has no Julia.  This file is
# the thin `ccall` mirror a maintainer of darchr/EmbeddingTables.jl would ship; every call below has a tested
# twin in the Python host mirror (embeddingtables.jl_b200/embtab), which drives exactly the same C ABI with the
# same arguments, and tests/abi_smoke.c drives that ABI from plain C.  Names, argument order and return
# conventions follow the reference (src/EmbeddingTables.jl:8-18 export list); tables live in HBM instead of
# Julia `Array`s.
#
# Provenance: the type shells a drop-in has to reproduce -- the `SimpleEmbedding` inner constructors
# (reference src/simple.jl:7-27), the shard loop of `SplitEmbedding` (src/split.jl:15-22, 29-46), the
# `PreallocationStrategy` constructors (src/lookup.jl:290-294) and the three lazy `rrule`s
# (src/lookup.jl:247-258, 374-389, src/sparseupdate.jl:35-40) -- are taken from the reference (MIT licence,
# Copyright (c) 2021 Mark Hildebrand) so that user code written against it keeps working; everything that
# computes is a `ccall` into hand-written sm_100a kernels.
#
# Execution model.  Every compute call is asynchronous on `stream()` (one non-blocking CUDA stream per
# process, created in `__init__`; `stream!(ptr)` installs the caller's own).  Nothing synchronises per call:
# `Array(::DeviceMatrix)`, scalar `getindex` and `synchronize()` are the only places that wait.  Index arrays
# are uploaded once (`DeviceIndices(I)`, done by the `rrule`s) and then reused by the forward pass, the lazy
# pullback and `update!`.
module EmbeddingTablesB200

export AbstractEmbeddingTable, SimpleEmbedding, SplitEmbedding, SparseEmbeddingUpdate, Static, Dynamic
export lookup, lookup!, maplookup, maplookup!, featuresize, example, columnpointer, columnview, update!, uncompress
export DefaultStrategy, SimpleParallelStrategy, PreallocationStrategy, Slicer, DeviceMatrix, DeviceIndices
export AbstractIndexer, Indexer, SparseIndexer, DenseIndexer, IndexerView, index!, ensemble_update
export IndexingContext, NoContext, Forward, Update, Descent, Adagrad
export synchronize, stream, stream!, set_update_order!

import ChainRulesCore: ChainRulesCore, NoTangent, ProjectTo
import Libdl

const libembtab = Ref{String}(get(ENV, "ETB_LIB_PATH", "libembtab_b200.so"))
const current_stream = Ref{Ptr{Cvoid}}(C_NULL)
stream() = current_stream[]
stream!(s::Ptr{Cvoid}) = (current_stream[] = s)                  # run on the caller's CUDA stream
synchronize() = check(ccall((:etb_stream_sync, libembtab[]), Int32, (Ptr{Cvoid},), stream()))

# ---------------------------------------------------------------------------------- C ABI structs
struct EtbTable              # etb_table
    base::Ptr{Cvoid}
    chunks::Ptr{Cvoid}
    nrows::Int64
    shard_rows::Int64
    dim::Int32
    ld::Int32
    elt::Int32
    reserved::Int32
end
struct EtbLookupItem         # etb_lookup_item
    table::EtbTable
    idx::Ptr{Cvoid}
    dst::Ptr{Cvoid}
    ld_dst::Int64
    batch::Int64
    bag::Int64
    ld_idx::Int64
    idx_elt::Int32
    reserved::Int32
end
struct EtbUpdateItem         # etb_update_item
    table::EtbTable
    delta::Ptr{Cvoid}
    ld_delta::Int64
    idx::Ptr{Cvoid}
    batch::Int64
    bag::Int64
    ld_idx::Int64
    idx_elt::Int32
    flags::Int32             # per-table ETB_UPDATE_FMA
end
mutable struct EtbIndexView  # etb_index_view (filled by etb_index)
    keys::Ptr{Cvoid}
    map::Ptr{Cvoid}
    records::Ptr{Cvoid}
    nnz::Ptr{Cvoid}
    scratch::Ptr{Cvoid}
    n_total::Int64
    key_bytes::Int32
    row_bits::Int32
    num_splits::Int32
    this_split::Int32
    EtbIndexView() = new(C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0, 0, 0, 0, 0)
end
const ETB_UPDATE_FMA = Int32(1)
const ETB_UPDATE_SPLIT_LONG = Int32(2)

etb_elt(::Type{Float32}) = Int32(0)
etb_elt(::Type{Float64}) = Int32(1)
etb_elt(::Type{Int32}) = Int32(2)
etb_elt(::Type{Int64}) = Int32(3)
etb_elt(::Type{Float16}) = Int32(4)   # extension: half-precision storage, Float32 arithmetic (ETB_F16)
# BFloat16s.BFloat16 maps to Int32(5) (ETB_BF16); add `etb_elt(::Type{BFloat16}) = Int32(5)` where BFloat16s is loaded

function check(status::Integer)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:etb_last_error, libembtab[]), Cstring, ()))
    error("libembtab_b200 status $status: $msg")
end

# Reduction order of update!.  :strict (default) adds the members of every bucket one after the other like the
# reference (bit-identical to it); :split sums buckets of more than 128 members as 128-member chunks combined in a
# fixed order (deterministic, within 1e-5 of the reference, a little faster on hot Zipf rows).
const update_order = Ref{Symbol}(:strict)
function set_update_order!(mode::Symbol)
    mode in (:strict, :split) || throw(ArgumentError("update order must be :strict or :split"))
    update_order[] = mode
end
order_flags() = update_order[] == :split ? ETB_UPDATE_SPLIT_LONG : Int32(0)

# ---------------------------------------------------------------------------------- device matrix
# Column-major matrix in HBM with a leading dimension: the `A` parameter of SimpleEmbedding{S,T,A}.
mutable struct DeviceMatrix{T} <: AbstractMatrix{T}
    ptr::Ptr{T}
    dims::Tuple{Int,Int}
    ld::Int
    owner::Any               # parent allocation (keeps it alive); `nothing` for the owner itself
end

function DeviceMatrix{T}(::UndefInitializer, m::Integer, n::Integer) where {T}
    p = Ref{Ptr{Cvoid}}()
    check(ccall((:etb_malloc, libembtab[]), Int32, (Ptr{Ptr{Cvoid}}, Csize_t), p, m * n * sizeof(T)))
    A = DeviceMatrix{T}(convert(Ptr{T}, p[]), (m, n), m, nothing)
    finalizer(x -> ccall((:etb_free, libembtab[]), Int32, (Ptr{Cvoid},), x.ptr), A)
    return A
end
function DeviceMatrix(h::Matrix{T}) where {T}
    A = DeviceMatrix{T}(undef, size(h)...)
    copyto!(A, h)
    return A
end
Base.size(A::DeviceMatrix) = A.dims
Base.strides(A::DeviceMatrix) = (1, A.ld)
Base.pointer(A::DeviceMatrix) = A.ptr
Base.similar(A::DeviceMatrix, ::Type{T}, dims::Tuple{Int,Int}) where {T} = DeviceMatrix{T}(undef, dims...)
Base.similar(A::DeviceMatrix{T}) where {T} = DeviceMatrix{T}(undef, size(A)...)
# view(A, rows, :) and view(A, :, cols) -- what PreallocationStrategy, its pullback and batch chunks need
rowview(A::DeviceMatrix{T}, r::UnitRange{Int}) where {T} =
    DeviceMatrix{T}(A.ptr + (first(r) - 1) * sizeof(T), (length(r), size(A, 2)), A.ld, A)
colview(A::DeviceMatrix{T}, c::UnitRange{Int}) where {T} =
    DeviceMatrix{T}(A.ptr + (first(c) - 1) * A.ld * sizeof(T), (size(A, 1), length(c)), A.ld, A)
Base.view(A::DeviceMatrix, r::UnitRange{Int}, ::Colon) = rowview(A, r)
Base.view(A::DeviceMatrix, ::Colon, c::UnitRange{Int}) = colview(A, c)
function Base.copyto!(A::DeviceMatrix{T}, h::Matrix{T}) where {T}
    size(A) == size(h) || throw(DimensionMismatch("copyto!: $(size(A)) vs $(size(h))"))
    # strided (2-D) copy: a row-slice view needs no staging
    check(ccall((:etb_memcpy2d_h2d, libembtab[]), Int32,
                (Ptr{Cvoid}, Csize_t, Ptr{Cvoid}, Csize_t, Csize_t, Csize_t, Ptr{Cvoid}),
                A.ptr, A.ld * sizeof(T), h, size(h, 1) * sizeof(T), size(h, 1) * sizeof(T), size(h, 2), stream()))
    synchronize()            # `h` is pageable Julia memory: it may be freed or changed right after this call
    return A
end
function Base.Array(A::DeviceMatrix{T}) where {T}
    h = Matrix{T}(undef, size(A)...)
    check(ccall((:etb_memcpy2d_d2h, libembtab[]), Int32,
                (Ptr{Cvoid}, Csize_t, Ptr{Cvoid}, Csize_t, Csize_t, Csize_t, Ptr{Cvoid}),
                h, size(h, 1) * sizeof(T), A.ptr, A.ld * sizeof(T), size(h, 1) * sizeof(T), size(h, 2), stream()))
    synchronize()
    return h
end
# scalar access: slow, tests only (reference src/EmbeddingTables.jl:144-156 does unsafe_load)
Base.getindex(A::DeviceMatrix, i::Int, j::Int) = Array(DeviceMatrix{eltype(A)}(
    A.ptr + ((j - 1) * A.ld + (i - 1)) * sizeof(eltype(A)), (1, 1), 1, A))[1]
function Base.setindex!(A::DeviceMatrix{T}, v, i::Int, j::Int) where {T}
    copyto!(DeviceMatrix{T}(A.ptr + ((j - 1) * A.ld + (i - 1)) * sizeof(T), (1, 1), 1, A), fill(convert(T, v), 1, 1))
    return v
end

# ---------------------------------------------------------------------------------- table types
abstract type AbstractExecutionStrategy end
abstract type AbstractLookupType end
struct Dynamic <: AbstractLookupType end
struct Static{N} <: AbstractLookupType end
Static(N) = Static{N}()
abstract type AbstractEmbeddingTable{S<:AbstractLookupType,T} <: AbstractArray{T,2} end
# access contexts (reference src/EmbeddingTables.jl:74-77): table types may place rows differently per phase
abstract type IndexingContext end
struct NoContext <: IndexingContext end
struct Forward <: IndexingContext end
struct Update <: IndexingContext end
featuresize(A::AbstractMatrix) = size(A, 1)
featuresize(::AbstractEmbeddingTable{Static{N}}) where {N} = N
featuresize(A::AbstractEmbeddingTable{Dynamic}) = size(A, 1)
example(x::AbstractVector{<:AbstractEmbeddingTable}) = example(first(x))
columnview(A::AbstractEmbeddingTable, i::Integer, ctx::IndexingContext = NoContext()) =
    DeviceMatrix{eltype(A)}(columnpointer(A, i, ctx), (featuresize(A), 1), featuresize(A), A)
Base.getindex(A::AbstractEmbeddingTable, i::Int, j::Int) = columnview(A, j)[i, 1]
Base.setindex!(A::AbstractEmbeddingTable, v, i::Int, j::Int) = (columnview(A, j)[i, 1] = v)

struct SimpleEmbedding{S,T,A<:AbstractMatrix{T}} <: AbstractEmbeddingTable{S,T}
    data::A
    SimpleEmbedding{Dynamic}(A::AbstractMatrix{T}) where {T} = new{Dynamic,T,typeof(A)}(A)
    SimpleEmbedding(A::AbstractMatrix) = SimpleEmbedding{Dynamic}(A)
    function SimpleEmbedding{Static{N}}(A::AbstractMatrix{T}) where {N,T}
        isa(N, Int) || throw(ArgumentError("Expected the type parameter for `Static{N}` to be an Int. Instead, it's a $(typeof(N))!"))
        N == size(A, 1) || throw(ArgumentError("Parameter `N` should match the number of rows in the passed Matrix. Instead, `N = $N` while `size(A,1) = $(size(A, 1))`."))
        return new{Static{N},T,typeof(A)}(A)
    end
end
SimpleEmbedding(A::AbstractMatrix, ::Val{N}) where {N} = SimpleEmbedding{Static{N}}(A)
Base.size(A::SimpleEmbedding) = size(A.data)
Base.parent(A::SimpleEmbedding) = A.data
Base.pointer(A::SimpleEmbedding) = pointer(A.data)
example(A::SimpleEmbedding) = A.data
# Dynamic honours the matrix's column stride (reference src/simple.jl:52, src/EmbeddingTables.jl:83-85);
# Static{N} assumes dense columns of N elements (src/simple.jl:53-55)
rowstride(A::SimpleEmbedding{Dynamic}) = strides(A.data)[2]
rowstride(::SimpleEmbedding{Static{N}}) where {N} = N
columnpointer(A::SimpleEmbedding{S,T}, i::Integer, ctx::IndexingContext = NoContext()) where {S,T} =
    pointer(A) + rowstride(A) * sizeof(T) * (i - 1)
function Base.zeros(x::SimpleEmbedding{S,T}) where {S,T}
    d = similar(x.data)
    check(ccall((:etb_memset, libembtab[]), Int32, (Ptr{Cvoid}, Int32, Csize_t, Ptr{Cvoid}), d.ptr, 0, sizeof(T) * length(d), stream()))
    return SimpleEmbedding{S}(d)
end
descriptor(A::SimpleEmbedding{S,T}) where {S,T} =
    (EtbTable(pointer(A), C_NULL, size(A, 2), 0, size(A, 1), rowstride(A), etb_elt(T), 0), nothing)

struct SplitEmbedding{S,T,A<:AbstractMatrix{T}} <: AbstractEmbeddingTable{S,T}
    data::Vector{A}
    matrixsize::Tuple{Int,Int}
    chunkptrs::DeviceMatrix{Int64}       # device array of chunk base pointers
end
chunk_pointer_array(data) = DeviceMatrix(reshape(Int64[Int64(UInt(pointer(d))) for d in data], :, 1))
function SplitEmbedding(A::Matrix{T}, cols_per_shard = 1) where {T}
    nshards = ceil(Int, size(A, 2) / cols_per_shard)
    data = map(1:nshards) do i
        start = cols_per_shard * (i - 1) + 1
        stop = min(i * cols_per_shard, size(A, 2))
        DeviceMatrix(A[:, start:stop])
    end
    return SplitEmbedding{Static{size(A, 1)},T,eltype(data)}(data, (size(A, 1), cols_per_shard), chunk_pointer_array(data))
end
# SplitEmbedding{S,T}(undef, featuresize, ncols, cols_per_shard) (reference src/split.jl:29-46)
__compare(::Type{Dynamic}, _) = nothing
__compare(::Type{Static{N}}, featuresize) where {N} = @assert N == featuresize
function SplitEmbedding{S,T}(::UndefInitializer, featuresize::Integer, ncols::Integer, cols_per_shard::Integer = 1) where {S,T}
    __compare(S, featuresize)
    nshards = ceil(Int, ncols / cols_per_shard)
    data = map(1:nshards) do i
        start = cols_per_shard * (i - 1) + 1
        stop = min(i * cols_per_shard, ncols)
        DeviceMatrix{T}(undef, featuresize, stop - start + 1)
    end
    return SplitEmbedding{S,T,eltype(data)}(data, (Int(featuresize), Int(cols_per_shard)), chunk_pointer_array(data))
end
Base.size(A::SplitEmbedding) = (A.matrixsize[1], A.matrixsize[2] * (length(A.data) - 1) + size(last(A.data), 2))
example(A::SplitEmbedding) = first(A.data)
function columnpointer(A::SplitEmbedding{S,T}, i::Integer, ctx::IndexingContext = NoContext()) where {S,T}
    chunk, col = divrem(i - 1, A.matrixsize[2])
    return pointer(A.data[chunk + 1]) + col * A.matrixsize[1] * sizeof(T)
end
descriptor(A::SplitEmbedding{S,T}) where {S,T} = (EtbTable(C_NULL, pointer(A.chunkptrs), size(A, 2), A.matrixsize[2],
                                                          A.matrixsize[1], A.matrixsize[1], etb_elt(T), 0), A.chunkptrs)

# ---------------------------------------------------------------------------------- indices
# Device-resident index arrays (Int64 like Julia's default, or Int32); a vector is stored n x 1 and flagged.
# Host arrays are uploaded ONCE by `DeviceIndices(I)`; lookup, the lazy pullback and update! then share that copy.
struct DeviceIndices{Ti<:Union{Int32,Int64}}
    data::DeviceMatrix{Ti}
    isvec::Bool
end
DeviceIndices(I::AbstractVector{Ti}) where {Ti<:Union{Int32,Int64}} = DeviceIndices{Ti}(DeviceMatrix(reshape(collect(I), :, 1)), true)
DeviceIndices(I::AbstractMatrix{Ti}) where {Ti<:Union{Int32,Int64}} = DeviceIndices{Ti}(DeviceMatrix(Matrix{Ti}(I)), false)
DeviceIndices(I::AbstractVecOrMat{<:Integer}) = DeviceIndices(Int64.(I))
_trailing_size(x::AbstractArray{<:Any,N}) where {N} = size(x, N)
_trailing_size(I::DeviceIndices) = I.isvec ? size(I.data, 1) : size(I.data, 2)
idx_elt(::DeviceIndices{Ti}) where {Ti} = etb_elt(Ti)
todevice(I::DeviceIndices) = I
todevice(I::AbstractArray{<:Integer}) = DeviceIndices(I)

# ---------------------------------------------------------------------------------- lookup
destination(A::AbstractEmbeddingTable, I) = similar(example(A), eltype(A), (featuresize(A), _trailing_size(I)))

function lookup_item(dst::DeviceMatrix, table::AbstractEmbeddingTable, I::DeviceIndices)
    desc, keep = descriptor(table)
    bag = I.isvec ? 0 : size(I.data, 1)
    return EtbLookupItem(desc, pointer(I.data), pointer(dst), dst.ld, _trailing_size(I), bag, I.isvec ? 0 : I.data.ld,
                         idx_elt(I), 0), keep
end

function run_lookup(items::Vector{EtbLookupItem})
    GC.@preserve items check(ccall((:etb_maplookup, libembtab[]), Int32, (Ptr{EtbLookupItem}, Int32, Ptr{Cvoid}),
                                   items, length(items), stream()))
end

function lookup!(dst, src::AbstractEmbeddingTable, indices)
    I = todevice(indices)
    item, keep = lookup_item(dst, src, I)
    GC.@preserve dst src I keep run_lookup([item])
    return dst
end
function lookup(A::AbstractEmbeddingTable, indices)
    I = todevice(indices)
    return lookup!(destination(A, I), A, I)
end

struct DefaultStrategy <: AbstractExecutionStrategy end
struct SimpleParallelStrategy <: AbstractExecutionStrategy end
struct PreallocationStrategy{T} <: AbstractExecutionStrategy
    prependrows::Int
end
PreallocationStrategy() = PreallocationStrategy{Any}(0)
PreallocationStrategy(x::Integer) = PreallocationStrategy{Any}(x)
# output element type: the strategy's `U` when given, else the tables' (reference src/lookup.jl:293-294, 312).  The
# kernels write the tables' element type, so `U` must agree with it.
_select_eltype(::Type{Any}, ::Type{T}) where {T} = T
_select_eltype(::Type{U}, ::Type{T}) where {U,T} =
    U === T ? T : throw(ArgumentError("PreallocationStrategy{$U}: the GPU path writes the tables' element type $T"))

colwrap(x::AbstractVector) = map(todevice, x)
colwrap(x::AbstractVector{<:DeviceIndices}) = x
colwrap(x::AbstractArray{<:Integer,N}) where {N} = [todevice(collect(selectdim(x, N, i))) for i in 1:size(x, N)]

maplookup(x::AbstractVector{<:AbstractEmbeddingTable}, I) = maplookup(DefaultStrategy(), x, I)
function maplookup(strategy::AbstractExecutionStrategy, x::AbstractVector{<:AbstractEmbeddingTable}, I0)
    I = colwrap(I0)
    return maplookup!(strategy, map(destination, x, I), x, I)
end
# Default and SimpleParallel: one fused launch writing the per-table outputs
function maplookup!(::Union{DefaultStrategy,SimpleParallelStrategy}, y::Vector, x::AbstractVector{<:AbstractEmbeddingTable}, I0)
    I = colwrap(I0)
    pairs = [lookup_item(y[i], x[i], I[i]) for i in eachindex(x)]
    GC.@preserve y x I pairs run_lookup([p[1] for p in pairs])
    return y
end
function maplookup(strategy::PreallocationStrategy{U}, x::Vector{<:AbstractEmbeddingTable{<:Any,T}}, I0; kw...) where {U,T}
    I = colwrap(I0)
    nrows = strategy.prependrows + sum(featuresize, x)
    dst = similar(example(x), _select_eltype(U, T), (nrows, _trailing_size(first(I))))
    return maplookup!(strategy, dst, x, I; kw...)
end
# Preallocation: the same launch, destinations = row blocks of the concatenated matrix; rows
# 1:prependrows are never touched (reference src/lookup.jl:311-313, 334-340)
function maplookup!(strategy::PreallocationStrategy, dst::DeviceMatrix, x::Vector{<:AbstractEmbeddingTable}, I0; worksize_div = 8)
    I = colwrap(I0)
    off = strategy.prependrows
    pairs = map(eachindex(x)) do i
        f = featuresize(x[i])
        p = lookup_item(rowview(dst, off+1:off+f), x[i], I[i])
        off += f
        p
    end
    GC.@preserve dst x I pairs run_lookup([p[1] for p in pairs])
    return dst
end

# ---------------------------------------------------------------------------------- sparse update
struct SparseEmbeddingUpdate{S<:AbstractLookupType,A,I}
    delta::A
    indices::I
end
SparseEmbeddingUpdate{S}(delta::A, indices::I) where {S,A,I} = SparseEmbeddingUpdate{S,A,I}(delta, indices)

# dense gradient of a sparse update (test helper, reference src/sparseupdate.jl:16-32)
function uncompress(x::SparseEmbeddingUpdate{S,<:DeviceMatrix{T}}, dstcols = maximum(Array(todevice(x.indices).data));
                    maxindices = size(x.delta, 2)) where {S,T}
    I = todevice(x.indices)
    dst = DeviceMatrix{T}(undef, size(x.delta, 1), dstcols)
    check(ccall((:etb_memset, libembtab[]), Int32, (Ptr{Cvoid}, Int32, Csize_t, Ptr{Cvoid}), dst.ptr, 0, sizeof(T) * length(dst), stream()))
    bag = I.isvec ? 0 : size(I.data, 1)
    GC.@preserve x I dst check(ccall((:etb_uncompress, libembtab[]), Int32,
        (Ptr{Cvoid}, Int64, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Ptr{Cvoid}),
        dst.ptr, dst.ld, size(x.delta, 1), etb_elt(T), pointer(x.delta), x.delta.ld, pointer(I.data), idx_elt(I), bag,
        min(size(x.delta, 2), maxindices), I.isvec ? 0 : I.data.ld, stream()))
    return dst
end

mutable struct Slicer{A}
    current_index::Int
    concat_dim::Int
    captured_array::A
end
function (S::Slicer)(sz)   # advances (the reference's copy does not; its test/map.jl:153-177 needs it to)
    r = S.current_index:(S.current_index + sz - 1)
    S.current_index += sz
    return rowview(S.captured_array, r)
end

# Caller-owned, reusable workspace in HBM (the GPU form of the reference's Indexer, src/utils.jl:288-304).  After
# index! it holds the sorted (row, delta column) pairs and one record per bucket (`view`).
abstract type AbstractIndexer end
mutable struct Indexer <: AbstractIndexer
    workspace::Union{Nothing,DeviceMatrix{UInt8}}
    view::EtbIndexView
    Indexer() = new(nothing, EtbIndexView())
end
const SparseIndexer = Indexer     # histogram flavours of the CPU algorithm (reference src/utils.jl:295-296);
const DenseIndexer = Indexer      # the GPU sort has one flavour
# IndexerView(I, num_splits, this_split) (reference src/utils.jl:320-333): a sub-range of the buckets
struct IndexerView <: AbstractIndexer
    I::Indexer
    num_splits::Int
    this_split::Int
end
ensemble_update(nthreads::Integer) = [Indexer() for _ in 1:nthreads]

# the reference's @generated dispatch (src/sparseupdate.jl:131-154): FMA epilogue for Static{N} Float32
# tables with N*4 <= 512 and N % 16 == 0 -- a per-table flag
update_flags(::AbstractEmbeddingTable{Static{N},Float32}) where {N} =
    (N * 4 <= 512 && N % 16 == 0) ? ETB_UPDATE_FMA : Int32(0)
update_flags(::AbstractEmbeddingTable) = Int32(0)

function update_item(table::AbstractEmbeddingTable, g::SparseEmbeddingUpdate)
    desc, keep = descriptor(table)
    I = todevice(g.indices)     # a DeviceIndices passes through: no upload here
    bag = I.isvec ? 0 : size(I.data, 1)
    return EtbUpdateItem(desc, pointer(g.delta), g.delta.ld, pointer(I.data), _trailing_size(I), bag,
                         I.isvec ? 0 : I.data.ld, idx_elt(I), update_flags(table)), (keep, I)
end

# index!(indexer, tables, grads): one batched sort for every table of the ensemble (the reference's
# `@batch index!` phase, src/sparseupdate.jl:211-213)
function index!(indexer::Indexer, tables, grads)
    pairs = [update_item(t, g) for (t, g) in zip(tables, grads)]
    items = [p[1] for p in pairs]
    need = Ref{Csize_t}(0)
    GC.@preserve items check(ccall((:etb_index_workspace_bytes, libembtab[]), Int32,
                                   (Ptr{EtbUpdateItem}, Int32, Ptr{Csize_t}), items, length(items), need))
    if indexer.workspace === nothing || length(indexer.workspace) < need[]
        indexer.workspace === nothing || synchronize()       # kernels may still read the old workspace
        indexer.workspace = DeviceMatrix{UInt8}(undef, need[], 1)
    end
    GC.@preserve items pairs tables grads check(ccall((:etb_index, libembtab[]), Int32,
        (Ptr{Cvoid}, Csize_t, Ptr{EtbUpdateItem}, Int32, Ref{EtbIndexView}, Ptr{Cvoid}),
        pointer(indexer.workspace), length(indexer.workspace), items, length(items), indexer.view, stream()))
    return indexer
end
index!(indexer::Indexer, table::AbstractEmbeddingTable, grad::SparseEmbeddingUpdate) = index!(indexer, [table], [grad])

# update!(table(s), update(s), indexer, alpha): apply an already-indexed update (reference src/sparseupdate.jl:46-154)
function apply!(tables, grads, indexer::AbstractIndexer, opt)
    pairs = [update_item(t, g) for (t, g) in zip(tables, grads)]
    items = [p[1] for p in pairs]
    base = indexer isa IndexerView ? indexer.I : indexer
    v = base.view
    view = EtbIndexView()
    view.keys, view.map, view.records, view.nnz, view.scratch = v.keys, v.map, v.records, v.nnz, v.scratch
    view.n_total, view.key_bytes, view.row_bits = v.n_total, v.key_bytes, v.row_bits
    if indexer isa IndexerView
        view.num_splits, view.this_split = indexer.num_splits, indexer.this_split
    end
    if opt isa Adagrad
        states = Ptr{Cvoid}[pointer(state!(opt, t)) for t in tables]
        GC.@preserve items pairs tables grads states check(ccall((:etb_adagrad_update, libembtab[]), Int32,
            (Ref{EtbIndexView}, Ptr{EtbUpdateItem}, Ptr{Ptr{Cvoid}}, Int32, Float64, Float64, Int32, Ptr{Cvoid}),
            view, items, states, length(items), Float64(opt.eta), Float64(opt.eps), order_flags(), stream()))
    else
        GC.@preserve items pairs tables grads check(ccall((:etb_sgd_update, libembtab[]), Int32,
            (Ref{EtbIndexView}, Ptr{EtbUpdateItem}, Int32, Float64, Int32, Ptr{Cvoid}),
            view, items, length(items), Float64(opt.eta), order_flags(), stream()))
    end
    return nothing
end
update!(table::AbstractEmbeddingTable, g::SparseEmbeddingUpdate, indexer::AbstractIndexer, alpha::Number, args...) =
    apply!([table], [g], indexer, Descent(alpha))

# Optimisers.  Any object with an `eta` field works as Flux.Descent (Flux stays a weak dependency, see ext/);
# `Descent` is the stand-in when Flux is not loaded.  `Adagrad` is a GPU-only extension (row-wise state).
struct Descent
    eta::Float64
end
Descent() = Descent(0.1)
mutable struct Adagrad
    eta::Float64
    eps::Float64
    state::IdDict{Any,Any}
end
Adagrad(eta = 0.1, eps = 1e-8) = Adagrad(eta, eps, IdDict{Any,Any}())
function state!(opt::Adagrad, t::AbstractEmbeddingTable{S,T}) where {S,T}
    get!(opt.state, t) do
        A = T === Float64 ? Float64 : Float32
        st = DeviceMatrix{A}(undef, size(t, 2), 1)
        check(ccall((:etb_memset, libembtab[]), Int32, (Ptr{Cvoid}, Int32, Csize_t, Ptr{Cvoid}), st.ptr, 0, sizeof(A) * length(st), stream()))
        st
    end
end

# update!(opt::Flux.Descent, table, grad, [indexer], [Val(nontemporal)]) (reference src/sparseupdate.jl:160-178):
# index! then the fused segment-reduce + SGD; returns nothing; asynchronous on stream()
function update!(opt, table::AbstractEmbeddingTable, g::SparseEmbeddingUpdate, indexer::Indexer = Indexer(), ::Val = Val(true), args...)
    index!(indexer, [table], [g])
    apply!([table], [g], indexer, opt)
    return nothing
end
# ensemble form (reference src/sparseupdate.jl:199-238); the CPU tuning keywords are accepted and ignored;
# telemetry_cb() runs between the index and the update phases like the reference's (:214)
function update!(opt, tables::AbstractVector{<:AbstractEmbeddingTable}, grads::AbstractVector{<:SparseEmbeddingUpdate},
                 indexers::AbstractVector{<:AbstractIndexer}, ::Val = Val(true); num_splits = 4, nthreads = 1,
                 scratchspaces = nothing, telemetry_cb = Returns(nothing))
    ix = first(indexers)
    index!(ix, tables, grads)
    telemetry_cb()
    apply!(tables, grads, ix, opt)
    return nothing
end

# ---------------------------------------------------------------------------------- rrules (lazy)
# the cotangent of a table is a SparseEmbeddingUpdate, not an array: projection passes it through
# (reference src/lookup.jl:246)
(p::ProjectTo{<:AbstractArray})(x::SparseEmbeddingUpdate) = x

function ChainRulesCore.rrule(::typeof(lookup), A::AbstractEmbeddingTable{S}, I) where {S}
    Id = todevice(I)          # uploaded once; forward, pullback and update! share it
    lookup_pullback(Δ) = (NoTangent(), SparseEmbeddingUpdate{S}(Δ, Id), NoTangent())
    return lookup(A, Id), lookup_pullback
end
function ChainRulesCore.rrule(::typeof(maplookup), strategy::AbstractExecutionStrategy,
                              A::Vector{<:AbstractEmbeddingTable{S}}, I) where {S}
    Is = colwrap(I)
    maplookup_pullback(Δs) = (NoTangent(), NoTangent(), map(SparseEmbeddingUpdate{S}, Δs, Is), NoTangent())
    return maplookup(strategy, A, Is), maplookup_pullback
end
function ChainRulesCore.rrule(::typeof(maplookup), strategy::PreallocationStrategy,
                              A::Vector{<:AbstractEmbeddingTable{S}}, I; kw...) where {S}
    Is = colwrap(I)
    data = maplookup(strategy, A, Is; kw...)
    function maplookup_pullback(Δ)
        f = Slicer(strategy.prependrows + 1, 1, Δ)
        δs = map((y, x) -> SparseEmbeddingUpdate{S}(f(featuresize(y)), x), A, Is)
        return (NoTangent(), NoTangent(), δs, NoTangent())
    end
    return data, maplookup_pullback
end

function __init__()
    Libdl.dlopen(libembtab[])
    check(ccall((:etb_init, libembtab[]), Int32, (Int32,), parse(Int32, get(ENV, "ETB_DEVICE", "0"))))
    s = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:etb_stream_create, libembtab[]), Int32, (Ptr{Ptr{Cvoid}},), s))
    current_stream[] = s[]
end

end # module
