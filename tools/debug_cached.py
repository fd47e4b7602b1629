import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import embtab as E
import oracle as O
rng = np.random.default_rng(128)
dim, nrows, bag, batch, cache_rows = 128, 5000, 8, 512, int(os.environ.get("CAP", "300"))
base = rng.standard_normal((dim, nrows)).astype(np.float32)
cached = E.CachedEmbedding(base, cache_rows, E.Static(dim), min_count=2)
ref = O.Table(base.copy(order="F"), static=True)
opt = E.Descent(0.05)
w = 1.0 / np.arange(1, nrows + 1) ** 1.05; cdf = np.cumsum(w); cdf /= cdf[-1]
for step in range(3):
    I = (rng.permutation(nrows)[np.searchsorted(cdf, rng.random(bag * batch))] + 1).reshape((bag, batch), order="F")
    out_c, back_c = E.pullback(E.lookup, cached, I)
    print("step", step, "fwd equal", np.array_equal(out_c.numpy(), O.lookup(ref, I)))
    delta = rng.standard_normal((dim, batch)).astype(np.float32)
    slots_before = cached._slot_of_row.cpu().numpy().copy()
    E.update_(opt, cached, back_c(delta)[1])
    O.update(ref, delta, I, 0.05)
    torch.cuda.synchronize()
    slots = cached._slot_of_row.cpu().numpy()
    g = E.lookup(cached, np.arange(1, nrows + 1)).numpy()
    bad = np.flatnonzero((g != ref.data).any(axis=0))
    cnt = np.bincount(I.reshape(-1) - 1, minlength=nrows)
    print("step", step, "cached rows", cached.cached_rows(), "bad rows", bad.size, "of which cached now", int((slots[bad] >= 0).sum()),
          "cached before", int((slots_before[bad] >= 0).sum()), "counts of bad", np.unique(cnt[bad])[:10], "host copy equal for bad:",
          [bool(np.array_equal(cached.host[:, r], ref.data[:, r])) for r in bad[:5]])
    if bad.size:
        r = bad[0]
        print(" row", r, "slot", slots[r], "got", g[:4, r], "want", ref.data[:4, r], "host", cached.host[:4, r], "base", base[:4, r])
