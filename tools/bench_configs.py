#!/usr/bin/env python
"""The other BASELINE.json configs (C1, C3, C5 sweep) on one B200.  Writes one JSON line per case to
stdout (and to --out).  bench.py stays the contract benchmark (C2); this script is the measurement of
SURVEY.md section 8d's remaining rows.

  python tools/bench_configs.py --config c1|c3|c5 [--out file.jsonl]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200")):
    sys.path.insert(0, p)
import torch

import bench
import embtab as E

PEAK, _ = bench.measured_peak_gbs()


def timeit(fn, iters=20, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)                      # write a buffer larger than L2 between iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return float(np.median(times))


def timeit_graph(fn, iters=10, warmup=3, flush=None):
    """GPU time of fn's launches alone: captured once into a CUDA graph and replayed, so the host-side
    enqueue cost of the Python mirror (~100 us per call) is not in the number."""
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    return timeit(g.replay, iters=iters, warmup=2, flush=flush)


def rand_tables(nt, dim, nrows, static=True):
    gen = torch.Generator(device="cuda").manual_seed(1)
    out = []
    for _ in range(nt):
        buf = torch.rand(dim * nrows, device="cuda", dtype=torch.float32, generator=gen)
        out.append(E.SimpleEmbedding(E.DeviceArray(buf, (dim, nrows)), E.Static(dim) if static else None))
    return out


def emit(rec, fh):
    line = json.dumps(rec)
    print(line, flush=True)
    if fh:
        fh.write(line + "\n")
        fh.flush()


def zipf_indices(rng, nrows, n, alpha=1.05):
    w = 1.0 / np.arange(1, nrows + 1, dtype=np.float64) ** alpha
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    ranks = np.searchsorted(cdf, rng.random(n))
    perm = rng.permutation(nrows)
    return (perm[ranks] + 1).astype(np.int64)


def config_c1(fh):
    """C1: 26 x (64 x 100k) f32, batch 2048, non-reducing maplookup + Descent update!.  27.7 MB forward /
    41 MB update of algorithmic traffic: launch-latency bound, so the step is also timed as one CUDA graph."""
    nt, dim, nrows, batch = 26, 64, 100_000, 2048
    rng = np.random.default_rng(0xE7AB1E + 1)
    tables = rand_tables(nt, dim, nrows)
    I = E.DeviceArray.from_numpy(rng.integers(1, nrows + 1, (batch, nt)))
    Is = list(E.colwrap(I))
    outs = [E.DeviceArray.empty((dim, batch)) for _ in range(nt)]
    deltas = [E.DeviceArray.from_numpy(rng.standard_normal((dim, batch)).astype(np.float32)) for _ in range(nt)]
    grads = [E.SparseEmbeddingUpdate(E.Static(dim), d, i) for d, i in zip(deltas, Is)]
    indexer, opt = E.Indexer(), E.Descent(0.01)
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")   # 256 MB > L2

    def fwd():
        E.maplookup_(E.DefaultStrategy(), outs, tables, I)

    def upd():
        E.update_(opt, tables, grads, [indexer])

    def step():
        fwd(); upd()

    t_fwd, t_upd, t_step = timeit(fwd, flush=flush), timeit(upd, flush=flush), timeit(step, flush=flush)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        step(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            step()
    t_graph = timeit(g.replay, flush=flush)

    def step_par():      # index! needs only the indices: on the side stream beside the gather (two branches in the graph)
        E.prefetch_index(indexer, tables, Is)
        fwd()
        E.update_(opt, tables, grads, [indexer])

    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        step_par(); torch.cuda.synchronize()
        with torch.cuda.graph(g2, stream=s):
            step_par()
    t_graph_par = timeit(g2.replay, flush=flush)
    u = sum(int(np.unique(i.numpy()).size) for i in Is)
    fwd_bytes = nt * batch * (8 + 2 * dim * 4)
    upd_bytes = nt * (batch * 8 + batch * dim * 4) + 2 * u * dim * 4
    emit({"config": "C1", "shape": f"{nt} x ({dim} x {nrows}) f32, batch {batch}, gather + update!",
          "fwd_us": t_fwd * 1e3, "update_us": t_upd * 1e3, "step_us": t_step * 1e3, "step_cuda_graph_us": t_graph * 1e3,
          "step_cuda_graph_index_beside_gather_us": t_graph_par * 1e3,
          "lookups_per_step": nt * batch, "lookups_per_sec_graph": nt * batch / (t_graph * 1e-3),
          "fwd_gbs": fwd_bytes / t_fwd / 1e6, "update_gbs": upd_bytes / t_upd / 1e6,
          "roofline_time_us_at_measured_peak": (fwd_bytes + upd_bytes) / PEAK / 1e3,
          "note": "L2 flushed between iterations; launch-bound (1 + 9 + 3 launches per step)"}, fh)


def config_c3(fh):
    """C3: one 128 x 10M f32 table, Zipf(1.05) indices: SparseEmbeddingUpdate + update!(Descent)."""
    dim, nrows = 128, 10_000_000
    rng = np.random.default_rng(0xE7AB1E + 3)
    table = rand_tables(1, dim, nrows)[0]
    indexer, opt = E.Indexer(), E.Descent(0.01)
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    cases = [("vector", 65536, 0), ("vector", 524288, 0), ("vector", 4194304, 0), ("pooled 32x16384", 524288, 32)]
    for name, n, bag in cases:
        idx = zipf_indices(rng, nrows, n)
        batch = n if bag == 0 else n // bag
        Ih = idx if bag == 0 else idx.reshape((bag, batch), order="F")
        I = E.DeviceArray.from_numpy(Ih)
        delta = E.DeviceArray(torch.randn(dim * batch, device="cuda"), (dim, batch))
        grad = E.SparseEmbeddingUpdate(E.Static(dim), delta, I)
        u = int(np.unique(idx).size)
        top = int(np.bincount(idx).max())
        bytes_alg = n * 8 + batch * dim * 4 + 2 * u * dim * 4
        for mode in ("split", "strict"):
            E.set_update_order(mode)
            t = timeit(lambda: E.update_(opt, table, grad, indexer), iters=10 if mode == "split" else 3,
                       warmup=2, flush=flush)
            ti = timeit(lambda: E.index_(indexer, table, grad), iters=10, warmup=2, flush=flush)
            tg = timeit_graph(lambda: E.update_(opt, table, grad, indexer), iters=10, warmup=2, flush=flush)   # GPU time alone
            tig = timeit_graph(lambda: E.index_(indexer, table, grad), iters=10, warmup=2, flush=flush)
            emit({"config": "C3", "form": name, "n": n, "distinct_rows": u, "hottest_row_members": top, "order": mode,
                  "update_us": t * 1e3, "index_us": ti * 1e3, "kernel_us": (t - ti) * 1e3,
                  "update_us_graph": tg * 1e3, "index_us_graph": tig * 1e3, "gbs_graph": bytes_alg / tg / 1e6,
                  "frac_of_measured_peak_graph": bytes_alg / tg / 1e6 / PEAK,
                  "lookups_per_sec": n / (t * 1e-3), "algorithmic_bytes": bytes_alg, "gbs": bytes_alg / t / 1e6,
                  "frac_of_measured_peak": bytes_alg / t / 1e6 / PEAK}, fh)
        E.set_update_order("strict")


def config_c5(fh, quick=False):
    """C5: pooled-lookup sweep, 26 tables x 1M rows, dim 16..256 x bag 1..128 x batch 1k..64k."""
    nrows, nt = 1_000_000, 26
    rng = np.random.default_rng(0xE7AB1E + 5)
    dims = [16, 32, 64, 128, 256]
    bags = [1, 2, 4, 8, 16, 32, 64, 128] if not quick else [1, 8, 32, 128]
    batches = [1024, 4096, 16384, 65536] if not quick else [4096, 16384]
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    for dim in dims:
        tables = rand_tables(nt, dim, nrows)
        for batch in batches:
            out = E.DeviceArray.empty((nt * dim, batch))
            for bag in bags:
                if nt * bag * batch * 8 > 12e9:
                    continue
                for dist in ("uniform", "zipf"):
                    if dist == "uniform":
                        I = E.DeviceArray(torch.randint(1, nrows + 1, (bag * batch * nt,), device="cuda", dtype=torch.int64),
                                          (bag, batch, nt))
                    else:
                        if batch * bag > 2 ** 21:
                            continue
                        I = E.DeviceArray.from_numpy(np.stack(
                            [zipf_indices(rng, nrows, bag * batch).reshape((bag, batch), order="F") for _ in range(nt)], axis=2))
                    fn = lambda: E.maplookup_(E.PreallocationStrategy(0), out, tables, I)
                    t = timeit_graph(fn, iters=10, warmup=3, flush=flush if dist == "uniform" else None)
                    b = nt * batch * (bag * (8 + dim * 4) + dim * 4)
                    emit({"config": "C5", "dim": dim, "bag": bag, "batch": batch, "dist": dist, "us": t * 1e3,
                          "lookups_per_sec": nt * batch * bag / (t * 1e-3), "gbs": b / t / 1e6,
                          "frac_of_measured_peak": b / t / 1e6 / PEAK, "timing": "CUDA-graph replay, L2 flushed (uniform)"}, fh)
                    del I
        del tables
        torch.cuda.empty_cache()


def config_c4local(fh):
    """C4's per-GPU share on one GPU: 8 SplitEmbedding tables 128 x 5M f32 (cols_per_shard 1 048 576: 5 chunks,
    last ragged), bag 32, global batch 131 072: fused pooled lookup into the (8*128) x B matrix + ensemble update!.
    (The exchange itself is measured by bench.py --gpus 8 on C2; this isolates chunked-table addressing.)"""
    nt, dim, nrows, shard, bag = 8, 128, 5_000_000, 1_048_576, 32
    gen = torch.Generator(device="cuda").manual_seed(4)
    for kind in ("split", "simple"):
        tables = []
        for _ in range(nt):
            if kind == "split":
                chunks = [E.DeviceArray(torch.rand(dim * (min(s + shard, nrows) - s), device="cuda", generator=gen),
                                        (dim, min(s + shard, nrows) - s)) for s in range(0, nrows, shard)]
                tables.append(E.SplitEmbedding(None, shard, _chunks=chunks, _lookup_type=E.Static(dim), _dtype=np.float32, _fs=dim))
            else:
                tables.append(E.SimpleEmbedding(E.DeviceArray(torch.rand(dim * nrows, device="cuda", generator=gen), (dim, nrows)), E.Static(dim)))
        for batch in (16384, 131072):
            I = E.DeviceArray(torch.randint(1, nrows + 1, (bag * batch * nt,), device="cuda", dtype=torch.int64), (bag, batch, nt))
            Is = list(E.colwrap(I))
            out = E.DeviceArray.empty((nt * dim, batch))
            delta = E.DeviceArray(torch.randn(nt * dim * batch, device="cuda"), (nt * dim, batch))
            grads = [E.SparseEmbeddingUpdate(E.Static(dim), delta.rows(k * dim, (k + 1) * dim), i) for k, i in enumerate(Is)]
            indexer, opt = E.Indexer(), E.Descent(0.01)
            t_f = timeit_graph(lambda: E.maplookup_(E.PreallocationStrategy(0), out, tables, I), iters=10, warmup=2)
            t_u = timeit(lambda: E.update_(opt, tables, grads, [indexer]), iters=10, warmup=2)
            fb = nt * batch * (bag * (8 + dim * 4) + dim * 4)
            emit({"config": "C4-local", "tables": kind, "shape": f"{nt} x (128 x 5M), bag {bag}, batch {batch}",
                  "fwd_ms": t_f, "fwd_gbs": fb / t_f / 1e6, "fwd_frac_of_measured_peak": fb / t_f / 1e6 / PEAK,
                  "fwd_lookups_per_sec": nt * batch * bag / (t_f * 1e-3), "update_ms": t_u,
                  "step_lookups_per_sec": nt * batch * bag / ((t_f + t_u) * 1e-3)}, fh)
            del I, out, delta, grads, indexer
        del tables
        torch.cuda.empty_cache()


def config_update_sweep(fh):
    """update!(Descent) over feature sizes: 26 tables x 1M rows, bag 32, batch 16384, uniform indices (C2's shape
    with dim varied) -- checks that the update path has no cliffs away from dim 128."""
    nrows, nt, bag, batch = 1_000_000, 26, 32, 16384
    rng = np.random.default_rng(0xE7AB1E + 6)
    I_host = rng.integers(1, nrows + 1, (bag, batch, nt))
    u = sum(int(np.unique(I_host[:, :, t]).size) for t in range(nt))
    I = E.DeviceArray.from_numpy(I_host)
    Is = list(E.colwrap(I))
    for dim in (16, 32, 64, 80, 128, 256, 512):
        tables = rand_tables(nt, dim, nrows)
        delta = E.DeviceArray(torch.randn(nt * dim * batch, device="cuda"), (nt * dim, batch))
        grads = [E.SparseEmbeddingUpdate(E.Static(dim), delta.rows(k * dim, (k + 1) * dim), i) for k, i in enumerate(Is)]
        indexer, opt = E.Indexer(), E.Descent(0.01)
        E.index_(indexer, tables, grads)
        t_k = timeit(lambda: E.sparseupdate._apply(tables, grads, indexer, 0.01), iters=10, warmup=3)
        t_i = timeit(lambda: E.index_(indexer, tables, grads), iters=10, warmup=3)
        b = nt * batch * dim * 4 + 2 * u * dim * 4 + nt * batch * bag * 4
        emit({"config": "update-sweep", "dim": dim, "kernel_ms": t_k, "index_ms": t_i, "algorithmic_bytes": b,
              "kernel_gbs": b / t_k / 1e6, "kernel_frac_of_measured_peak": b / t_k / 1e6 / PEAK,
              "lookups_per_sec_update": nt * batch * bag / ((t_k + t_i) * 1e-3)}, fh)
        del tables, delta, grads, indexer
        torch.cuda.empty_cache()


def config_lowp(fh):
    """C2's shape with half-precision tables (ETB_F16 / ETB_BF16, Float32 arithmetic) beside Float32: the forward
    and the update kernels move half the bytes per row."""
    nrows, nt, bag, batch, dim = 1_000_000, 26, 32, 16384, 128
    rng = np.random.default_rng(0xE7AB1E + 7)
    I_host = rng.integers(1, nrows + 1, (bag, batch, nt))
    u = sum(int(np.unique(I_host[:, :, t]).size) for t in range(nt))
    I = E.DeviceArray.from_numpy(I_host)
    Is = list(E.colwrap(I))
    for name, tdt, ndt in (("f32", torch.float32, np.float32), ("f16", torch.float16, np.float16), ("bf16", torch.bfloat16, E.bfloat16)):
        es = np.dtype(ndt).itemsize
        gen = torch.Generator(device="cuda").manual_seed(1)
        tables = []
        for _ in range(nt):
            buf = torch.rand(dim * nrows, device="cuda", dtype=torch.float32, generator=gen).to(tdt)
            tables.append(E.SimpleEmbedding(E.DeviceArray(buf, (dim, nrows)), E.Static(dim)))
        out = E.DeviceArray.empty((128 + nt * dim, batch), ndt)
        strategy = E.PreallocationStrategy(128)
        t_f = timeit_graph(lambda: E.maplookup_(strategy, out, tables, I), iters=10, warmup=3)   # GPU time (the eager call is host-bound here)
        delta = E.DeviceArray(torch.randn(nt * dim * batch, device="cuda").to(tdt), (nt * dim, batch))
        grads = [E.SparseEmbeddingUpdate(E.Static(dim), delta.rows(k * dim, (k + 1) * dim), i) for k, i in enumerate(Is)]
        indexer = E.Indexer()
        E.index_(indexer, tables, grads)
        t_k = timeit_graph(lambda: E.sparseupdate._apply(tables, grads, indexer, 0.01), iters=10, warmup=3)
        t_i = timeit_graph(lambda: E.index_(indexer, tables, grads), iters=10, warmup=3)
        fb = nt * batch * (bag * (8 + dim * es) + dim * es)
        ub = nt * batch * dim * es + 2 * u * dim * es + nt * batch * bag * 4
        step = t_f + t_i + t_k
        opt = E.Adagrad(0.01, 1e-8)
        for t in tables:
            opt.state(t)                                     # allocate the state vectors outside the timed region
        t_a = timeit_graph(lambda: E.sparseupdate._apply(tables, grads, indexer, 0.01, opt), iters=10, warmup=3)
        ab = ub + 2 * u * 4                                   # + state read-modify-write, one Float32 per distinct row
        emit({"config": "c2-adagrad", "dtype": name, "update_ms": t_a, "sgd_update_ms": t_k, "update_gbs": ab / t_a / 1e6,
              "update_frac_of_measured_peak": ab / t_a / 1e6 / PEAK}, fh)
        emit({"config": "c2-lowp", "dtype": name, "fwd_ms": t_f, "index_ms": t_i, "update_ms": t_k, "step_ms_back_to_back": step,
              "lookups_per_sec": nt * batch * bag / (step * 1e-3), "fwd_gbs": fb / t_f / 1e6, "fwd_frac_of_measured_peak": fb / t_f / 1e6 / PEAK,
              "update_gbs": ub / t_k / 1e6, "update_frac_of_measured_peak": ub / t_k / 1e6 / PEAK}, fh)
        del tables, delta, grads, indexer, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c1", "c3", "c4local", "c5", "c5quick", "update", "lowp"])
    ap.add_argument("--out")
    a = ap.parse_args()
    E._lib.check(E.lib().etb_init(0))
    fh = open(a.out, "a") if a.out else None
    {"lowp": config_lowp, "c1": config_c1, "c3": config_c3, "c4local": config_c4local, "update": config_update_sweep, "c5": config_c5, "c5quick": lambda f: config_c5(f, True)}[a.config](fh)
