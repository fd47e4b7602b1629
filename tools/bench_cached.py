#!/usr/bin/env python
"""Host-tier table behind an HBM row cache (CachedEmbedding) under Zipf(1.05): hit rate and throughput
of the pooled lookup and of update!(Descent) as the cache warms up, beside the all-HBM table.  One JSON line per case.

  python tools/bench_cached.py [--rows 1000000] [--fractions 0.01,0.1,0.25] [--out file.jsonl]
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
from bench_configs import timeit

PEAK, _ = bench.measured_peak_gbs()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--bag", type=int, default=32)
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--fractions", default="0.01,0.1,0.25")
    ap.add_argument("--warm-steps", type=int, default=4)
    ap.add_argument("--out")
    a = ap.parse_args()
    E._lib.check(E.lib().etb_init(0))
    rng = np.random.default_rng(0xCAC4E)
    dim, nrows, bag, batch = a.dim, a.rows, a.bag, a.batch
    base = rng.random((dim, nrows), dtype=np.float32)
    fh = open(a.out, "a") if a.out else None
    n = bag * batch
    w = 1.0 / np.arange(1, nrows + 1, dtype=np.float64) ** 1.05
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    perm = rng.permutation(nrows)          # ONE rank -> row map: the hot rows stay hot from batch to batch
    batches = [(perm[np.searchsorted(cdf, rng.random(n))] + 1).reshape((bag, batch), order="F") for _ in range(a.warm_steps + 1)]
    delta = E.DeviceArray(torch.randn(dim * batch, device="cuda"), (dim, batch))
    opt = E.Descent(0.01)
    fwd_bytes = batch * (bag * (8 + dim * 4) + dim * 4)

    def measure(table, name, capacity):
        out = E.DeviceArray.empty((dim, batch))
        ix = E.Indexer()
        hits = []
        for I in batches[:-1]:                                  # warm-up steps fill the cache (Update-phase hook)
            Id = E.as_device_indices(I)
            hits.append(table.hit_rate(I) if capacity is not None else 1.0)
            E.lookup_(out, table, Id)
            E.update_(opt, table, E.SparseEmbeddingUpdate(table.lookup_type, delta, Id), ix)
        I = batches[-1]
        Id = E.as_device_indices(I)
        hit = table.hit_rate(I) if capacity is not None else 1.0
        t_f = timeit(lambda: E.lookup_(out, table, Id), iters=5, warmup=1)
        t_u = timeit(lambda: E.update_(opt, table, E.SparseEmbeddingUpdate(table.lookup_type, delta, Id), ix), iters=3, warmup=1)
        u = int(np.unique(I).size)
        upd_bytes = n * 8 + batch * dim * 4 + 2 * u * dim * 4
        rec = {"config": "cached-table", "table": name, "rows": nrows, "dim": dim, "bag": bag, "batch": batch,
               "dist": "zipf(1.05)", "cache_rows": capacity, "cached_rows": table.cached_rows() if capacity is not None else None,
               "hit_rate_per_warm_step": hits, "hit_rate": hit, "fwd_ms": t_f, "fwd_lookups_per_sec": n / (t_f * 1e-3),
               "fwd_gbs": fwd_bytes / t_f / 1e6, "fwd_frac_of_measured_hbm_peak": fwd_bytes / t_f / 1e6 / PEAK,
               "fwd_pcie_gbs": (1 - hit) * n * dim * 4 / t_f / 1e6, "update_ms": t_u, "update_gbs": upd_bytes / t_u / 1e6}
        line = json.dumps(rec)
        print(line, flush=True)
        if fh:
            fh.write(line + "\n")
            fh.flush()

    measure(E.SimpleEmbedding(base.copy(), E.Static(dim)), "SimpleEmbedding (all in HBM)", None)
    for f in [float(x) for x in a.fractions.split(",")]:
        cap = int(nrows * f)
        measure(E.CachedEmbedding(base, cap, E.Static(dim), min_count=2), f"CachedEmbedding ({f:.0%} of the rows in HBM)", cap)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
