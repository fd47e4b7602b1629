"""Non-reducing gather throughput (K1) at C5-like sizes: 26 tables x 1M rows, batch 16384/65536, GPU time by graph replay."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import embtab as E
from bench_configs import rand_tables, timeit_graph, PEAK
nt, nrows = 26, 1_000_000
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
for dim in (16, 64, 128, 256):
    tables = rand_tables(nt, dim, nrows)
    for batch in (16384, 65536):
        I = E.DeviceArray(torch.randint(1, nrows + 1, (batch * nt,), device="cuda", dtype=torch.int64), (batch, nt))
        outs = [E.DeviceArray.empty((dim, batch)) for _ in range(nt)]
        t = timeit_graph(lambda: E.maplookup_(E.DefaultStrategy(), outs, tables, I), iters=10, warmup=3, flush=flush)
        b = nt * batch * (8 + 2 * dim * 4)
        print(json.dumps({"config": "gather", "dim": dim, "batch": batch, "us": t * 1e3, "gbs": b / t / 1e6, "frac_of_measured_peak": b / t / 1e6 / PEAK}), flush=True)
    del tables
    torch.cuda.empty_cache()
