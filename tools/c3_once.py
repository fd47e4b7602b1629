import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import embtab as E
from bench_configs import zipf_indices, rand_tables
dim, nrows, n = 128, 10_000_000, 524288
rng = np.random.default_rng(3)
table = rand_tables(1, dim, nrows)[0]
I = E.DeviceArray.from_numpy(zipf_indices(rng, nrows, n))
delta = E.DeviceArray(torch.randn(dim * n, device="cuda"), (dim, n))
grad = E.SparseEmbeddingUpdate(E.Static(dim), delta, I)
ix, opt = E.Indexer(), E.Descent(0.01)
E.set_update_order(os.environ.get("ETB_ORDER", "strict"))
for _ in range(4):
    E.update_(opt, table, grad, ix)
torch.cuda.synchronize()
print("done")
