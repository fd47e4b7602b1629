# Flux compatibility (reference src/sparseupdate.jl:180-189): `Flux.Optimise.update!(opt, table, grad, ...)` with
# a SparseEmbeddingUpdate as the gradient forwards to this package's update!, so `Flux.Descent(eta)` (any optimiser
# object with an `eta` field) works unchanged in user code.  Loaded automatically when Flux is (Julia >= 1.9
# package extension; on older Julia `include` this file after `using Flux`).
# NOT EXECUTED IN THIS REPOSITORY'S ENVIRONMENT (no Julia in the image).
module EmbeddingTablesB200FluxExt

import Flux
using EmbeddingTablesB200: EmbeddingTablesB200, AbstractEmbeddingTable, SparseEmbeddingUpdate, Indexer, update!

function Flux.Optimise.update!(opt, x::AbstractEmbeddingTable, xbar::SparseEmbeddingUpdate, indexer = Indexer(),
                               nontemporal::Val = Val(true), args...)
    return update!(opt, x, xbar, indexer, nontemporal, args...)
end

end # module
