# EmbeddingTablesB200 -- Julia host layer over libembtab_b200.so (include/embtab_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S ENVIRONMENT: the build image has no Julia.  This file is the
# thin `ccall` mirror a maintainer of darchr/EmbeddingTables.jl would ship; every call below has a
# tested twin in the Python host mirror (embeddingtables.jl_b200/embtab), which drives exactly the same
# C ABI with the same arguments.  Names, argument order and return conventions follow the reference
# (src/EmbeddingTables.jl:8-18 export list); tables live in HBM instead of Julia `Array`s.
module EmbeddingTablesB200

export AbstractEmbeddingTable, SimpleEmbedding, SplitEmbedding, SparseEmbeddingUpdate, Static, Dynamic
export lookup, lookup!, maplookup, maplookup!, featuresize, example, columnpointer, update!
export DefaultStrategy, SimpleParallelStrategy, PreallocationStrategy, Slicer, Indexer, DeviceMatrix

import ChainRulesCore: ChainRulesCore, NoTangent
import Libdl

const libembtab = Ref{String}(get(ENV, "ETB_LIB_PATH", "libembtab_b200.so"))

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
# view(A, rows, :) -- what PreallocationStrategy and its pullback need
rowview(A::DeviceMatrix{T}, r::UnitRange{Int}) where {T} =
    DeviceMatrix{T}(A.ptr + (first(r) - 1) * sizeof(T), (length(r), size(A, 2)), A.ld, A)
function Base.copyto!(A::DeviceMatrix{T}, h::Matrix{T}) where {T}
    @assert A.ld == size(A, 1) && size(A) == size(h)
    check(ccall((:etb_memcpy_h2d, libembtab[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                A.ptr, h, sizeof(h), C_NULL))
    check(ccall((:etb_stream_sync, libembtab[]), Int32, (Ptr{Cvoid},), C_NULL))
    return A
end
function Base.Array(A::DeviceMatrix{T}) where {T}
    @assert A.ld == size(A, 1)
    h = Matrix{T}(undef, size(A)...)
    check(ccall((:etb_memcpy_d2h, libembtab[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                h, A.ptr, sizeof(h), C_NULL))
    check(ccall((:etb_stream_sync, libembtab[]), Int32, (Ptr{Cvoid},), C_NULL))
    return h
end
# scalar access: slow, tests only (reference src/EmbeddingTables.jl:144-156 does unsafe_load)
Base.getindex(A::DeviceMatrix, i::Int, j::Int) = Array(DeviceMatrix{eltype(A)}(
    A.ptr + ((j - 1) * A.ld + (i - 1)) * sizeof(eltype(A)), (1, 1), 1, A))[1]

# ---------------------------------------------------------------------------------- table types
abstract type AbstractExecutionStrategy end
abstract type AbstractLookupType end
struct Dynamic <: AbstractLookupType end
struct Static{N} <: AbstractLookupType end
Static(N) = Static{N}()
abstract type AbstractEmbeddingTable{S<:AbstractLookupType,T} <: AbstractArray{T,2} end
featuresize(A::AbstractMatrix) = size(A, 1)
featuresize(::AbstractEmbeddingTable{Static{N}}) where {N} = N
featuresize(A::AbstractEmbeddingTable{Dynamic}) = size(A, 1)
example(x::AbstractVector{<:AbstractEmbeddingTable}) = example(first(x))

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
Base.size(A::SimpleEmbedding) = size(A.data)
Base.parent(A::SimpleEmbedding) = A.data
Base.pointer(A::SimpleEmbedding) = pointer(A.data)
example(A::SimpleEmbedding) = A.data
columnpointer(A::SimpleEmbedding{S,T}, i::Integer, ctx...) where {S,T} =
    pointer(A) + strides(A.data)[2] * sizeof(T) * (i - 1)
Base.zeros(x::SimpleEmbedding{S,T}) where {S,T} = (d = similar(x.data);
    check(ccall((:etb_memset, libembtab[]), Int32, (Ptr{Cvoid}, Int32, Csize_t, Ptr{Cvoid}), d.ptr, 0, sizeof(T) * length(d), C_NULL));
    SimpleEmbedding{S}(d))
descriptor(A::SimpleEmbedding{S,T}) where {S,T} = (EtbTable(pointer(A), C_NULL, size(A, 2), 0, size(A, 1),
                                                           strides(A.data)[2], etb_elt(T), 0), nothing)

struct SplitEmbedding{S,T,A<:AbstractMatrix{T}} <: AbstractEmbeddingTable{S,T}
    data::Vector{A}
    matrixsize::Tuple{Int,Int}
    chunkptrs::DeviceMatrix{Int64}       # device array of chunk base pointers
end
function SplitEmbedding(A::Matrix{T}, cols_per_shard = 1) where {T}
    nshards = ceil(Int, size(A, 2) / cols_per_shard)
    data = map(1:nshards) do i
        start = cols_per_shard * (i - 1) + 1
        stop = min(i * cols_per_shard, size(A, 2))
        DeviceMatrix(A[:, start:stop])
    end
    ptrs = DeviceMatrix(reshape(Int64[Int64(UInt(pointer(d))) for d in data], :, 1))
    return SplitEmbedding{Static{size(A, 1)},T,eltype(data)}(data, (size(A, 1), cols_per_shard), ptrs)
end
Base.size(A::SplitEmbedding) = (A.matrixsize[1], A.matrixsize[2] * (length(A.data) - 1) + size(last(A.data), 2))
example(A::SplitEmbedding) = first(A.data)
function columnpointer(A::SplitEmbedding{S,T}, i::Integer, ctx...) where {S,T}
    chunk, col = divrem(i - 1, A.matrixsize[2])
    return pointer(A.data[chunk + 1]) + col * A.matrixsize[1] * sizeof(T)
end
descriptor(A::SplitEmbedding{S,T}) where {S,T} = (EtbTable(C_NULL, pointer(A.chunkptrs), size(A, 2), A.matrixsize[2],
                                                          A.matrixsize[1], A.matrixsize[1], etb_elt(T), 0), A.chunkptrs)

# ---------------------------------------------------------------------------------- lookup
_trailing_size(x::AbstractArray{<:Any,N}) where {N} = size(x, N)
destination(A::AbstractEmbeddingTable, I) = similar(example(A), eltype(A), (featuresize(A), _trailing_size(I)))

# Device-resident index arrays are DeviceMatrix{Int64} (vectors are n x 1 with bag = 0 semantics
# selected by `isvec`); host `Array`s are uploaded first.
struct DeviceIndices
    data::DeviceMatrix{Int64}
    isvec::Bool
end
DeviceIndices(I::AbstractVector{<:Integer}) = DeviceIndices(DeviceMatrix(reshape(Int64.(I), :, 1)), true)
DeviceIndices(I::AbstractMatrix{<:Integer}) = DeviceIndices(DeviceMatrix(Matrix{Int64}(I)), false)
_trailing_size(I::DeviceIndices) = I.isvec ? size(I.data, 1) : size(I.data, 2)
todevice(I::DeviceIndices) = I
todevice(I::AbstractArray{<:Integer}) = DeviceIndices(I)

function lookup_item(dst::DeviceMatrix, table::AbstractEmbeddingTable, I::DeviceIndices)
    desc, keep = descriptor(table)
    bag = I.isvec ? 0 : size(I.data, 1)
    return EtbLookupItem(desc, pointer(I.data), pointer(dst), dst.ld, _trailing_size(I), bag, I.isvec ? 0 : I.data.ld,
                         etb_elt(Int64), 0), keep
end

function run_lookup(items::Vector{EtbLookupItem})
    GC.@preserve items check(ccall((:etb_maplookup, libembtab[]), Int32, (Ptr{EtbLookupItem}, Int32, Ptr{Cvoid}),
                                   items, length(items), C_NULL))
end

function lookup!(dst, src::AbstractEmbeddingTable, indices)
    I = todevice(indices)
    item, keep = lookup_item(dst, src, I)
    GC.@preserve dst src I keep run_lookup([item])
    return dst
end
lookup(A::AbstractEmbeddingTable, I) = lookup!(destination(A, todevice(I)), A, todevice(I))

struct DefaultStrategy <: AbstractExecutionStrategy end
struct SimpleParallelStrategy <: AbstractExecutionStrategy end
struct PreallocationStrategy{T} <: AbstractExecutionStrategy
    prependrows::Int
end
PreallocationStrategy() = PreallocationStrategy{Any}(0)
PreallocationStrategy(x::Integer) = PreallocationStrategy{Any}(x)

colwrap(x::AbstractVector) = map(todevice, x)
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
function maplookup(strategy::PreallocationStrategy, x::Vector{<:AbstractEmbeddingTable{<:Any,T}}, I0; kw...) where {T}
    I = colwrap(I0)
    nrows = strategy.prependrows + sum(featuresize, x)
    dst = similar(example(x), T, (nrows, _trailing_size(first(I))))
    return maplookup!(strategy, dst, x, I)
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

# Caller-owned, reusable workspace in HBM (the GPU form of the reference's Indexer)
mutable struct Indexer
    workspace::Union{Nothing,DeviceMatrix{UInt8}}
    Indexer() = new(nothing)
end

# the reference's @generated dispatch (src/sparseupdate.jl:131-154): FMA epilogue for Static{N} Float32
# tables with N*4 <= 512 and N % 16 == 0 -- a per-table flag
update_flags(::AbstractEmbeddingTable{Static{N},Float32}) where {N} =
    (N * 4 <= 512 && N % 16 == 0) ? ETB_UPDATE_FMA : Int32(0)
update_flags(::AbstractEmbeddingTable) = Int32(0)

function update_item(table::AbstractEmbeddingTable, g::SparseEmbeddingUpdate)
    desc, keep = descriptor(table)
    I = todevice(g.indices)
    bag = I.isvec ? 0 : size(I.data, 1)
    return EtbUpdateItem(desc, pointer(g.delta), g.delta.ld, pointer(I.data), _trailing_size(I), bag,
                         I.isvec ? 0 : I.data.ld, etb_elt(Int64), update_flags(table)), (keep, I)
end

function update_many!(eta, tables, grads, indexer::Indexer)
    pairs = [update_item(t, g) for (t, g) in zip(tables, grads)]
    items = [p[1] for p in pairs]
    need = Ref{Csize_t}(0)
    GC.@preserve items check(ccall((:etb_index_workspace_bytes, libembtab[]), Int32,
                                   (Ptr{EtbUpdateItem}, Int32, Ptr{Csize_t}), items, length(items), need))
    if indexer.workspace === nothing || length(indexer.workspace) < need[]
        indexer.workspace = DeviceMatrix{UInt8}(undef, need[], 1)
    end
    GC.@preserve items pairs tables grads check(ccall((:etb_index_and_update, libembtab[]), Int32,
        (Ptr{Cvoid}, Csize_t, Ptr{EtbUpdateItem}, Int32, Float64, Int32, Ptr{Cvoid}),
        pointer(indexer.workspace), length(indexer.workspace), items, length(items), Float64(eta),
        ETB_UPDATE_SPLIT_LONG, C_NULL))
    check(ccall((:etb_stream_sync, libembtab[]), Int32, (Ptr{Cvoid},), C_NULL))
    return nothing
end

# update!(opt::Flux.Descent, table, grad, [indexer], [Val(nontemporal)]): any optimiser with an
# `eta` field is accepted so that Flux stays a weak dependency (reference src/sparseupdate.jl:160-178)
update!(opt, table::AbstractEmbeddingTable, g::SparseEmbeddingUpdate, indexer::Indexer = Indexer(), ::Val = Val(true), args...) =
    update_many!(opt.eta, [table], [g], indexer)
# ensemble form (reference src/sparseupdate.jl:199-238); CPU tuning keywords accepted and ignored
function update!(opt, tables::AbstractVector{<:AbstractEmbeddingTable}, grads::AbstractVector{<:SparseEmbeddingUpdate},
                 indexers::AbstractVector{Indexer}, ::Val = Val(true); num_splits = 4, nthreads = 1,
                 scratchspaces = nothing, telemetry_cb = Returns(nothing))
    update_many!(opt.eta, tables, grads, first(indexers))
    telemetry_cb()
    return nothing
end

# ---------------------------------------------------------------------------------- rrules (lazy)
function ChainRulesCore.rrule(::typeof(lookup), A::AbstractEmbeddingTable{S}, I) where {S}
    Id = todevice(I)
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
    check(ccall((:etb_init, libembtab[]), Int32, (Int32,), 0))
end

end # module
