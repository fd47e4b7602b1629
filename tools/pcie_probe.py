#!/usr/bin/env python
"""What the box's host <-> device path delivers when N ranks use it at once (the floor of bench.py's e2e at N > 1):
every rank copies a 256 MB pinned buffer H2D, D2H and both at once; per-rank and aggregate GB/s, max time over ranks.

  python tools/pcie_probe.py                                                    # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29730 tools/pcie_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

MB = 256


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = MB * 1024 * 1024
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a, d_b = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, iters=10):
        fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sync()
        return t.item()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d(); d2h()

    res = {}
    for name, fn, bytes_ in (("h2d", h2d, n), ("d2h", d2h, n), ("h2d+d2h", both, 2 * n)):
        ms = timed(fn)
        res[name] = {"ms_max_over_ranks": ms, "gbs_per_rank": bytes_ / ms / 1e6, "gbs_aggregate": world * bytes_ / ms / 1e6}
    if rank == 0:
        print(json.dumps({"probe": "host<->device copies from pinned memory, all ranks at once", "n_gpus": world,
                          "buffer_mb": MB, **res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
