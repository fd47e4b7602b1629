"""CUDA-graph capture of a step (GPU-only extension).

Every C-ABI call is allocation-free, never synchronises and takes its descriptors by value in kernel
parameters, so a whole forward + index! + update! sequence can be captured once and replayed with one
launch.  That is what makes tiny configurations (BASELINE C1: 27 MB of traffic, ~10 us at roofline)
launch-latency- instead of host-bound: 298 us eager -> 86 us replayed on B200.

    step = embtab.capture(lambda: (embtab.maplookup_(strategy, out, tables, I),
                                   embtab.update_(opt, tables, grads, [indexer])))
    step()          # replays the captured launches on the current stream

The captured callable must not allocate device memory: pass preallocated outputs (`maplookup_`,
`lookup_`) and a warmed-up `Indexer` (the warm-up calls below size its workspace).
"""
from __future__ import annotations

import torch


def capture(fn, warmup: int = 2):
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(1, warmup)):
            fn()                      # sizes workspaces, loads kernels
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)

    def replay():
        graph.replay()

    replay.graph = graph
    return replay
