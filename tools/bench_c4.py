#!/usr/bin/env python
"""BASELINE configs[3] ("C4") under torchrun: SplitEmbedding (chunked) tables 128 x 5M Float32, cols_per_shard
1 048 576 (5 chunks, the last ragged -- reference src/split.jl:3-86), 8 tables per GPU block-partitioned table-wise
(64 tables on 8 GPUs), bag 32, pooled lookup + backward, global batch 16 384 and 131 072, uniform and Zipf(1.05).
One step = fused lookup storing into the peers' feature matrices over NVLink, the reverse exchange of the cotangent and
the owners' ensemble update!.  Device-timed, max over ranks; one JSON line per case on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 \
      tools/bench_c4.py [--batches 16384,131072] [--dists uniform,zipf] [--out file.jsonl]
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
import torch.distributed as dist

import bench
import embtab as E
from embtab.dist import ShardedEnsemble, ShardPlan

TABLES_PER_GPU, DIM, NROWS, SHARD, BAG, PREPEND = 8, 128, 5_000_000, 1_048_576, 32, 128


def zipf_on_device(nrows, n, gen, alpha=1.05):
    """Zipf(alpha) ranks by inverse CDF on the GPU (the CDF over 5M rows is 40 MB), rank -> row by a fixed permutation"""
    w = 1.0 / torch.arange(1, nrows + 1, device="cuda", dtype=torch.float64) ** alpha
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    ranks = torch.searchsorted(cdf, torch.rand(n, device="cuda", dtype=torch.float64, generator=gen)).clamp_(max=nrows - 1)
    perm = torch.randperm(nrows, device="cuda", generator=gen)
    return perm[ranks] + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="16384,131072")
    ap.add_argument("--dists", default="uniform,zipf")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--groups", type=int, default=int(os.environ.get("ETB_TABLE_GROUPS", "1")))
    ap.add_argument("--out")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    E._lib.check(E.lib().etb_init(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peak, _ = bench.measured_peak_gbs()
    gen = torch.Generator(device="cuda").manual_seed(0xC4 + rank)
    tables = []
    for _ in range(TABLES_PER_GPU):
        chunks = [E.DeviceArray(torch.rand(DIM * (min(s + SHARD, NROWS) - s), device="cuda", generator=gen),
                                (DIM, min(s + SHARD, NROWS) - s)) for s in range(0, NROWS, SHARD)]
        tables.append(E.SplitEmbedding(None, SHARD, _chunks=chunks, _lookup_type=E.Static(DIM), _dtype=np.float32, _fs=DIM))
    fh = open(a.out, "a") if (a.out and rank == 0) else None
    opt = E.Descent(0.01)
    for batch in [int(b) for b in a.batches.split(",")]:
        plan = ShardPlan([DIM] * (TABLES_PER_GPU * world), world, rank, PREPEND, batch)
        ens = ShardedEnsemble(tables, plan, fused=True, table_groups=a.groups)
        for dname in a.dists.split(","):
            n = BAG * batch
            if dname == "uniform":
                I = torch.randint(1, NROWS + 1, (n * TABLES_PER_GPU,), device="cuda", dtype=torch.int64, generator=gen)
            else:
                I = torch.cat([zipf_on_device(NROWS, n, gen) for _ in range(TABLES_PER_GPU)])
            I_dev = E.DeviceArray(I, (BAG, batch, TABLES_PER_GPU))
            delta = E.DeviceArray(torch.randn(plan.total_rows * plan.my_cols, device="cuda", generator=gen),
                                  (plan.total_rows, plan.my_cols))

            def step(ev=None):
                ens.forward(I_dev)
                if ev: ev[1].record()
                ens.backward_update_(opt, delta)
                if ev: ev[2].record()

            for _ in range(a.warmup):
                step()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
            for k in range(a.steps):
                ev[k][0].record()
                step(ev[k])
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([ev[0][0].elapsed_time(ev[-1][2]) / a.steps,
                              float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
                              float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, fwd_ms, bwd_ms = t.tolist()
            # the backward exchange alone (scatter of every group + barriers), for the NVLink figure
            for g in range(ens.n_groups):
                ens.scatter_group(delta, g)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                for g in range(ens.n_groups):
                    ens.scatter_group(delta, g)
            e1.record()
            torch.cuda.synchronize()
            x = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
            if rank == 0:
                lookups = world * TABLES_PER_GPU * batch * BAG
                fwd_bytes = TABLES_PER_GPU * batch * (BAG * (8 + DIM * 4) + DIM * 4)      # per GPU, SURVEY 8d pooled fwd
                sent = plan.my_rows * batch * 4 * (world - 1) / world                       # per GPU and direction
                rec = {"config": "C4", "n_gpus": world, "tables": f"{TABLES_PER_GPU * world} SplitEmbedding 128 x 5M f32, cols_per_shard {SHARD}",
                       "batch_global": batch, "bag": BAG, "dist": dname, "table_groups": ens.n_groups,
                       "ms_per_step": ms, "fwd_lookup+exchange_ms": fwd_ms, "bwd_exchange+update_ms": bwd_ms,
                       "lookups_per_sec": lookups / (ms * 1e-3),
                       "fwd_gbs_per_gpu": fwd_bytes / fwd_ms / 1e6, "fwd_frac_of_measured_hbm_peak": fwd_bytes / fwd_ms / 1e6 / peak,
                       "a2a_bytes_sent_per_gpu_per_direction": sent, "bwd_exchange_alone_ms": x.item(),
                       "bwd_exchange_gbs_per_gpu": sent / x.item() / 1e6 if world > 1 else None,
                       "nvlink_peak_gbs": 900.0,
                       "bwd_exchange_frac_of_nvlink": sent / x.item() / 1e6 / 900.0 if world > 1 else None}
                line = json.dumps(rec)
                print(line, flush=True)
                if fh:
                    fh.write(line + "\n"); fh.flush()
            del I, I_dev, delta
            torch.cuda.empty_cache()
        ens.close()
        del ens
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
