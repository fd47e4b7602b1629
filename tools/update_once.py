"""C2's update shape with the feature size varied (python tools/update_once.py DIM): a few update! calls for a profiler."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import embtab as E
from bench_configs import rand_tables
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nrows, nt, bag, batch = 1_000_000, 26, 32, 16384
rng = np.random.default_rng(0xE7AB1E + 6)
I = E.DeviceArray.from_numpy(rng.integers(1, nrows + 1, (bag, batch, nt)))
Is = list(E.colwrap(I))
tables = rand_tables(nt, dim, nrows)
delta = E.DeviceArray(torch.randn(nt * dim * batch, device="cuda"), (nt * dim, batch))
grads = [E.SparseEmbeddingUpdate(E.Static(dim), delta.rows(k * dim, (k + 1) * dim), i) for k, i in enumerate(Is)]
indexer = E.Indexer()
E.index_(indexer, tables, grads)
for _ in range(4):
    E.sparseupdate._apply(tables, grads, indexer, 0.01)
torch.cuda.synchronize()
print("done")
