# Mirror of the reference's test strategy (test/lookup.jl, test/map.jl, test/update.jl) against the B200
# library, with a plain-Matrix CPU statement of the same definitions as the checker.
# NOT EXECUTED IN THIS REPOSITORY'S ENVIRONMENT (no Julia in the image): the executed twin of every test
# below is tests/test_gpu_lookup.py / tests/test_gpu_update.py, which drive the same C ABI.
using EmbeddingTablesB200, Test, Random
import ChainRulesCore

# dense definitions (reference src/lookup.jl:6-13 with the table path's sequential bag order)
ref_lookup(A::Matrix, I::AbstractVector) = A[:, I]
function ref_lookup(A::Matrix, I::AbstractMatrix)
    O = A[:, I[1, :]]
    for i in 2:size(I, 1)
        O .+= A[:, I[i, :]]
    end
    return O
end

@testset "lookup" begin
    for rows in [32, 64, 128, 256, 512, 1024, 1504]          # reference test/lookup.jl:67
        base = rand(Float32, rows, 1000)
        for table in (SimpleEmbedding(DeviceMatrix(base)), SimpleEmbedding{Static{rows}}(DeviceMatrix(base)),
                      SplitEmbedding(base, 30))
            @test size(table) == size(base)
            I = rand(1:1000, 1000)
            @test Array(lookup(table, I)) == ref_lookup(base, I)
            I = rand(1:1000, 12, 999)
            @test Array(lookup(table, I)) == ref_lookup(base, I)
        end
    end
end

@testset "maplookup" begin                                   # reference test/map.jl:14-100
    for nrows in [16, 64, 512]
        base = [randn(Float32, nrows, 100) for _ in 1:10]
        tables = [SimpleEmbedding{Static{nrows}}(DeviceMatrix(b)) for b in base]
        for inds in ([rand(1:100, 64) for _ in 1:10], rand(1:100, 64, 10),
                     [rand(1:100, 10, 64) for _ in 1:10], rand(1:100, 10, 64, 10))
            Is = inds isa Vector ? inds : [collect(selectdim(inds, ndims(inds), i)) for i in 1:10]
            reference = reduce(vcat, map(ref_lookup, base, Is))
            @test reduce(vcat, map(Array, maplookup(DefaultStrategy(), tables, inds))) == reference
            @test reduce(vcat, map(Array, maplookup(SimpleParallelStrategy(), tables, inds))) == reference
            @test Array(maplookup(PreallocationStrategy(), tables, inds)) == reference
            @test Array(maplookup(PreallocationStrategy(20), tables, inds))[21:end, :] == reference
        end
    end
end

@testset "update" begin                                      # reference test/update.jl:4-84
    for rows in [64, 80, 256], reducing in (false, true)
        base = randn(Float32, rows, 100)
        table = SimpleEmbedding{Static{rows}}(DeviceMatrix(copy(base)))
        I = reducing ? rand(1:100, 10, 100) : rand(1:100, 100)
        out, back = ChainRulesCore.rrule(lookup, table, I)
        @test Array(out) == ref_lookup(base, I)
        Δ = randn(Float32, size(out)...)
        g = back(DeviceMatrix(Δ))[2]
        @test g isa SparseEmbeddingUpdate
        update!(Descent(10.0), table, g)
        dense = zeros(Float32, size(base))
        flat = vec(I)
        bag = reducing ? size(I, 1) : 1
        for (p, c) in enumerate(flat)
            dense[:, c] .+= Δ[:, div(p - 1, bag) + 1]
        end
        @test isapprox(Array(parent(table)), base .- 10.0f0 .* dense)     # reference tolerance (isapprox)
    end
end

@testset "ensemble update, IndexerView, uncompress" begin     # reference test/update.jl:64-120, src/sparseupdate.jl:16-32
    base = [randn(Float32, 64, 100) for _ in 1:4]
    tables = [SimpleEmbedding{Static{64}}(DeviceMatrix(copy(b))) for b in base]
    Is = [rand(1:100, 10, 50) for _ in 1:4]
    out, back = ChainRulesCore.rrule(maplookup, PreallocationStrategy(8), tables, Is)
    Δ = randn(Float32, size(out)...)
    grads = back(DeviceMatrix(Δ))[3]
    called = Ref(0)
    update!(Descent(0.5), tables, grads, [Indexer()]; telemetry_cb = () -> (called[] += 1))
    @test called[] == 1
    for t in 1:4
        dense = zeros(Float32, 64, 100)
        for (p, c) in enumerate(vec(Is[t]))
            dense[:, c] .+= Δ[8 + 64 * (t - 1) .+ (1:64), div(p - 1, 10) + 1]
        end
        @test isapprox(Array(parent(tables[t])), base[t] .- 0.5f0 .* dense)
        @test isapprox(Array(uncompress(grads[t], 100)), dense)
    end
    # partitioned update == full update (reference test/update.jl:90-118)
    a = SimpleEmbedding{Static{64}}(DeviceMatrix(copy(base[1])))
    b = SimpleEmbedding{Static{64}}(DeviceMatrix(copy(base[1])))
    g = SparseEmbeddingUpdate{Static{64}}(DeviceMatrix(Δ[9:72, :]), DeviceIndices(Is[1]))
    ix = Indexer()
    update!(Descent(0.5), a, g, ix)
    index!(ix, b, g)
    for j in 1:4
        update!(b, g, IndexerView(ix, 4, j), 0.5)
    end
    @test Array(parent(a)) == Array(parent(b))
    # undef constructor (reference src/split.jl:29-46)
    s = SplitEmbedding{Static{16},Float32}(undef, 16, 1000, 300)
    @test size(s) == (16, 1000)
end
