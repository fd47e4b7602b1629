#!/usr/bin/env python
"""index! (K4) alone: GPU time by CUDA-graph replay at the BASELINE shapes, plus a check of the buckets against
numpy's stable sort.  One JSON line per case.

  python tools/index_bench.py [--cases c2,c2zipf,c3,c1,c4] [--out file.jsonl]
  ETB_IX_THREADS=512 python tools/index_bench.py ...   # the other tile size of the partition passes
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch

import bench
import embtab as E
from bench_configs import timeit_graph, zipf_indices
from embtab.sparseupdate import _IndicesOnly, _peek

CASES = {
    # name: (tables, rows, bag, batch, dist)
    "c2": (26, 1_000_000, 32, 16384, "uniform"),
    "c2zipf": (26, 1_000_000, 32, 16384, "zipf"),
    "c3": (1, 10_000_000, 0, 524288, "zipf_raw"),
    "c3pooled": (1, 10_000_000, 32, 16384, "zipf_raw"),
    "c1": (26, 100_000, 0, 2048, "uniform"),
    "c4": (8, 5_000_000, 32, 16384, "uniform"),
}


class Declared(E.SimpleEmbedding):
    """index! only looks at the declared row count"""

    def __init__(self, nrows):
        super().__init__(np.zeros((4, 8), np.float32))
        self._n = nrows

    def descriptor(self):
        d = super().descriptor()
        d.nrows = self._n
        return d


def run(name, fh):
    nt, nrows, bag, batch, dist = CASES[name]
    rng = np.random.default_rng(7)
    n = batch * max(bag, 1)
    Is = []
    for _ in range(nt):
        if dist == "uniform":
            i = rng.integers(1, nrows + 1, n)
        elif dist == "zipf":
            i = zipf_indices(rng, nrows, n)
        else:  # C3: rank = row (no permutation), the hottest row first
            w = 1.0 / np.arange(1, nrows + 1, dtype=np.float64) ** 1.05
            cdf = np.cumsum(w)
            cdf /= cdf[-1]
            i = np.searchsorted(cdf, rng.random(n)) + 1
        Is.append(np.asfortranarray(i.reshape((bag, batch) if bag else (batch,), order="F")).astype(np.int64))
    tables = [Declared(nrows) for _ in range(nt)]
    dI = [E.as_device_indices(i) for i in Is]
    grads = [_IndicesOnly(t, i) for t, i in zip(tables, dI)]
    ix = E.Indexer()
    E.index_(ix, tables, grads)
    torch.cuda.synchronize()
    launches = E.lib().etb_last_launch_count()
    # ---- check against numpy's stable sort, table by table
    v = ix.view
    keys = _peek(v.keys, v.n_total, np.uint32 if v.key_bytes == 4 else np.uint64).astype(np.int64)
    mp = _peek(v.map, v.n_total, np.int32)
    nnz = int(_peek(v.nnz, 1, np.int64)[0])
    rec = _peek(v.records, nnz, np.dtype([("start", np.uint32), ("m0", np.int32), ("key", np.uint64)]))
    ok, off, want_nnz, heads = True, 0, 0, []
    for t, i in enumerate(Is):
        flat = i.reshape(-1, order="F")
        order = np.argsort(flat, kind="stable")
        cols = (order // bag if bag else order).astype(np.int32)
        ok &= np.array_equal(keys[off:off + n], flat[order] - 1) and np.array_equal(mp[off:off + n], cols)
        srt = flat[order]
        h = np.flatnonzero(np.r_[True, srt[1:] != srt[:-1]])
        heads.append((off + h, cols[h], (np.uint64(t) << np.uint64(v.row_bits)) | (srt[h] - 1).astype(np.uint64)))
        want_nnz += h.size
        off += n
    ok &= nnz == want_nnz
    if ok:
        ok &= np.array_equal(rec["start"], np.concatenate([h[0] for h in heads]).astype(np.uint32))
        ok &= np.array_equal(rec["m0"], np.concatenate([h[1] for h in heads]))
        ok &= np.array_equal(rec["key"], np.concatenate([h[2] for h in heads]))
    ms = timeit_graph(lambda: E.index_(ix, tables, grads), iters=30)
    rec_out = {"case": name, "tables": nt, "rows": nrows, "n_per_table": n, "dist": dist, "index_ms": ms,
               "launches": launches, "keys_per_s": nt * n / (ms * 1e-3), "ok": bool(ok), "nnz": nnz,
               "ix_threads": os.environ.get("ETB_IX_THREADS", "default"), "row_bits": int(v.row_bits)}
    line = json.dumps(rec_out)
    print(line, flush=True)
    if fh:
        fh.write(line + "\n")
        fh.flush()
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c2,c2zipf,c3,c3pooled,c1,c4")
    ap.add_argument("--out")
    a = ap.parse_args()
    E._lib.check(E.lib().etb_init(0))
    fh = open(a.out, "a") if a.out else None
    good = True
    for c in a.cases.split(","):
        good &= run(c, fh)
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
