"""Does the C2 end-to-end step get faster when the kernels read/write the HOST buffers themselves?

The staged e2e step of bench.py copies the feature matrix out (D2H, 226 MB) and the cotangent in (H2D, 226 MB)
with the copy engines, around the kernels.  Page-locked host memory is addressable from the device (UVA), so the
forward kernel can store its result straight into the pinned feature matrix and the update kernel can load the
cotangent columns straight from the pinned cotangent: the PCIe transfer then happens INSIDE the kernels, overlapped
with their HBM traffic.  This script measures both kernels in that mode, checks the results against the staged
path bit for bit, and times the whole zero-copy e2e step.  Output: JSON lines (gpurun_out/zc_probe.jsonl).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch

import embtab as E
from bench import BAG, BATCH, DIM, ETA, NROWS, NT, PREPEND, SEED, make_indices

OUT = os.path.join(ROOT, "gpurun_out", "zc_probe.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def emit(**kw):
    s = json.dumps(kw)
    print(s, flush=True)
    with open(OUT, "a") as f:
        f.write(s + "\n")


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / iters
    return max(a.elapsed_time(b) / iters, wall)


def main():
    torch.cuda.set_device(0)
    E._lib.check(E.lib().etb_init(0))
    rng = np.random.default_rng(SEED)
    gen = torch.Generator(device="cuda").manual_seed(SEED)
    bufs = [torch.rand(DIM * NROWS, device="cuda", dtype=torch.float32, generator=gen) for _ in range(NT)]
    mk = lambda bs: [E.SimpleEmbedding(E.DeviceArray(b, (DIM, NROWS)), E.Static(DIM)) for b in bs]
    tables = mk(bufs)
    I_host = make_indices(rng, "uniform", NT, NROWS, BAG, BATCH)
    idx_pinned = E.pinned_empty((BAG, BATCH, NT), np.int64)
    idx_pinned[...] = I_host
    total_rows = PREPEND + NT * DIM
    out_pinned = E.pinned_empty((total_rows, BATCH), np.float32)
    out_pinned[...] = 0
    delta_pinned = E.pinned_empty((total_rows, BATCH), np.float32)
    delta_pinned.reshape(-1, order="F")[:] = rng.standard_normal(total_rows * BATCH, dtype=np.float32)
    I_dev = E.DeviceArray.empty((BAG, BATCH, NT), np.int64).upload(idx_pinned)
    Is = list(E.colwrap(I_dev))
    out_dev = E.DeviceArray.empty((total_rows, BATCH), np.float32)
    out_dev.buf.zero_()
    delta_dev = E.DeviceArray.empty((total_rows, BATCH), np.float32).upload(delta_pinned)
    strategy = E.PreallocationStrategy(PREPEND)
    S = E.Static(DIM)
    opt = E.Descent(ETA)
    indexer = E.Indexer()

    # device views of the pinned host buffers
    out_zc, delta_zc = E.DeviceArray.mapped(out_pinned), E.DeviceArray.mapped(delta_pinned)

    # ---- forward: staged vs zero-copy ---------------------------------------------------------------
    E.maplookup_(strategy, out_dev, tables, I_dev)
    E.maplookup_(strategy, out_zc, tables, I_dev)
    torch.cuda.synchronize()
    ref = np.empty((total_rows, BATCH), np.float32, order="F")
    out_dev.download(ref)
    torch.cuda.synchronize()
    same = bool(np.array_equal(ref[PREPEND:].view(np.uint32), out_pinned[PREPEND:].view(np.uint32)))
    t_dev = timed(lambda: E.maplookup_(strategy, out_dev, tables, I_dev))
    t_zc = timed(lambda: E.maplookup_(strategy, out_zc, tables, I_dev))
    t_copy = timed(lambda: out_dev.download(out_pinned))
    mb = out_pinned.nbytes / 1e6
    emit(stage="forward", bit_identical=same, device_ms=t_dev, zero_copy_ms=t_zc, d2h_copy_ms=t_copy,
         zero_copy_gbs=mb / t_zc, d2h_copy_gbs=mb / t_copy)

    # ---- update: cotangent in HBM vs cotangent read from the host by the kernel ------------------------
    def grads_of(d):
        slicer = E.Slicer(PREPEND + 1, 1, d)
        return [E.SparseEmbeddingUpdate(S, slicer(DIM), i) for i in Is]
    tables_b = mk([b.clone() for b in bufs])
    E.update_(opt, tables, grads_of(delta_dev), [indexer])
    E.update_(opt, tables_b, grads_of(delta_zc), [indexer])
    torch.cuda.synchronize()
    same = all(torch.equal(a.data.buf, b.data.buf) for a, b in zip(tables, tables_b))
    del tables_b
    torch.cuda.empty_cache()
    g_dev, g_zc = grads_of(delta_dev), grads_of(delta_zc)
    E.index_(indexer, tables, g_dev)
    t_dev = timed(lambda: E.sparseupdate._apply(tables, g_dev, indexer, opt.eta))
    t_zc = timed(lambda: E.sparseupdate._apply(tables, g_zc, indexer, opt.eta), iters=3, warmup=1)
    t_copy = timed(lambda: delta_dev.upload(delta_pinned))
    emit(stage="update", bit_identical=bool(same), device_ms=t_dev, zero_copy_ms=t_zc, h2d_copy_ms=t_copy,
         zero_copy_gbs=mb / t_zc, h2d_copy_gbs=mb / t_copy)

    # ---- whole e2e step, zero-copy: indices by copy engine (double-buffered), result and cotangent by the kernels --
    copy_stream = torch.cuda.Stream()
    I_buf = [I_dev, E.DeviceArray.empty((BAG, BATCH, NT), np.int64)]
    Is_buf = [Is, list(E.colwrap(I_buf[1]))]
    idx_ready, buf_free, st = [None, None], [None, None], {"k": 0}

    def upload_indices(slot):
        with torch.cuda.stream(copy_stream):
            if buf_free[slot] is not None:
                copy_stream.wait_event(buf_free[slot])
            I_buf[slot].upload(idx_pinned)
            idx_ready[slot] = copy_stream.record_event()

    def make_step(fwd_zc, upd_zc):
        def step():
            k = st["k"]
            slot = k % 2
            if idx_ready[slot] is None:
                upload_indices(slot)
            main_s = torch.cuda.current_stream()
            main_s.wait_event(idx_ready[slot])
            idx_ready[slot] = None
            upload_indices(1 - slot)                   # next step's indices (H2D) while the result goes out (D2H)
            E.prefetch_index(indexer, tables, Is_buf[slot])
            if fwd_zc:
                E.maplookup_(strategy, out_zc, tables, Is_buf[slot])     # result lands in the pinned host matrix
            else:
                E.maplookup_(strategy, out_dev, tables, Is_buf[slot])
                out_dev.download(out_pinned)
            if not upd_zc:
                delta_dev.upload(delta_pinned)
            slicer = E.Slicer(PREPEND + 1, 1, delta_zc if upd_zc else delta_dev)
            grads = [E.SparseEmbeddingUpdate(S, slicer(DIM), i) for i in Is_buf[slot]]
            E.update_(opt, tables, grads, [indexer])
            buf_free[slot] = main_s.record_event()
            st["k"] = k + 1
        return step

    lookups = NT * BATCH * BAG
    for fwd_zc, upd_zc, iters in ((False, False, 10), (True, False, 10), (True, True, 3)):
        t = timed(make_step(fwd_zc, upd_zc), iters=iters, warmup=2)
        emit(stage="e2e", forward_stores_to_host=fwd_zc, update_loads_from_host=upd_zc, ms_per_step=t,
             lookups_per_sec=lookups / (t * 1e-3),
             h2d_bytes=int(idx_pinned.nbytes + delta_pinned.nbytes), d2h_bytes=int(out_pinned.nbytes))


if __name__ == "__main__":
    main()
