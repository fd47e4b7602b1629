"""embtab -- host-side mirror of EmbeddingTables.jl's hot path over libembtab_b200.so.

Same names and argument meaning as the reference's export list (src/EmbeddingTables.jl:8-18);
Julia's `f!` is spelled `f_`.  Device memory comes from PyTorch; every computation is a call
into the sm_100a C-ABI library.  There is no CPU fallback.
"""
from ._lib import EmbTabError, LIB_PATH, lib
from .cached import CachedEmbedding
from .darray import DeviceArray, as_device, as_device_indices, bfloat16, pinned_empty
from .graph import capture
from .lookup import (AbstractExecutionStrategy, ColumnWrap, DefaultStrategy, PreallocationStrategy,
                     SimpleParallelStrategy, colwrap, destination, lookup, lookup_, maplookup, maplookup_)
from .sparseupdate import (AbstractIndexer, Adagrad, DenseIndexer, Descent, Indexer, IndexerView, Slicer,
                           SparseEmbeddingUpdate, SparseIndexer, ensemble_update, index_, prefetch_index, pullback, rrule,
                           set_update_order,
                           uncompress, update_, update_table_)
from .tables import (AbstractEmbeddingTable, ArgumentError, Dynamic, Forward, IndexingContext, NoContext,
                     SimpleEmbedding, SplitEmbedding, Static, Update, columnpointer, example, featuresize,
                     zeros)

__all__ = [n for n in dir() if not n.startswith("_")]
