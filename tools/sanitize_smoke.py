"""Small pass over every kernel family for compute-sanitizer (memcheck): tiny shapes, odd dims, split
tables, long buckets, IndexerView, uncompress, pack/unpack/scatter.  Run:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200")):
    sys.path.insert(0, p)
import ctypes as C

import torch

import embtab as E
from embtab import _lib

rng = np.random.default_rng(0)
for dtype, dim in [(np.float32, 128), (np.float32, 5), (np.float64, 24), (np.int64, 16), (np.float32, 1504), (np.float32, 64)]:
    nrows = 333
    if np.issubdtype(dtype, np.integer):
        base = rng.integers(-1000, 1000, (dim, nrows)).astype(dtype)
    else:
        base = rng.standard_normal((dim, nrows)).astype(dtype)
    for table in (E.SimpleEmbedding(base.copy()), E.SplitEmbedding(base.copy(), 50)):
        for bag in (1, 2, 4, 7, 33):
            E.lookup(table, rng.integers(1, nrows + 1, (bag, 97)))
        E.lookup(table, rng.integers(1, nrows + 1, 131))
        E.lookup(table, rng.integers(1, nrows + 1, 1).astype(np.int32))
        if not np.issubdtype(dtype, np.integer):
            for mode in ("split", "strict"):
                E.set_update_order(mode)
                I = rng.integers(1, nrows + 1, (3, 700))
                I[:, :400] = 7                                      # one long bucket (1200 members)
                I[0, 400:500] = 9                                   # one medium bucket
                g = E.SparseEmbeddingUpdate(table.lookup_type, rng.standard_normal((dim, 700)).astype(dtype), I)
                ix = E.Indexer()
                E.update_(E.Descent(0.1), table, g, ix)
                E.index_(ix, table, g)
                for s in range(1, 4):
                    E.update_table_(table, g, E.IndexerView(ix, 3, s), 0.1)
                E.uncompress(g, nrows)
            E.set_update_order("strict")
tables = [E.SimpleEmbedding(rng.standard_normal((d, 90)).astype(np.float32)) for d in (16, 64, 5, 128)]
I = [rng.integers(1, 91, (4, 37)) for _ in tables]
out, back = E.pullback(E.maplookup, E.PreallocationStrategy(3), tables, I)
E.update_(E.Descent(0.1), tables, back(rng.standard_normal(out.shape).astype(np.float32))[2], [E.Indexer()])
# pack / unpack / scatter
rows = (C.c_int64 * 2)(16, 69)
offs = (C.c_int64 * 2)(3, 19)
dense = E.DeviceArray.zeros((85 * 37,))
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(_lib.lib().etb_a2a_pack(dense.ptr, out.ptr, out.ld, rows, offs, 2, 37, out.elt, stream))
_lib.check(_lib.lib().etb_a2a_unpack(out.ptr, out.ld, dense.ptr, rows, offs, 2, 37, out.elt, stream))
ptrs = (C.c_void_p * 2)(dense.ptr, dense.ptr + 16 * 37 * 4)
_lib.check(_lib.lib().etb_a2a_scatter(ptrs, out.ptr, out.ld, rows, offs, 2, 37, out.elt, stream))
torch.cuda.synchronize()
print("sanitize smoke done")
