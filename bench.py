#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dist uniform|zipf]

Workload (BASELINE.json configs[1], "C2"): DLRM-style pooled lookup, 26 tables x 1M rows x dim
128 Float32, bag 32, batch 16384, PreallocationStrategy(prependrows=128).  One STEP = the whole
hot path once over one batch: fused multi-table pooled lookup into the concatenated feature matrix
(forward), the lazy pullback (SparseEmbeddingUpdate views of the cotangent), and the ensemble
update!(Descent) (index! + fused segment-reduce + SGD).  metric = embedding lookups per second
(lookups per step / step time); fwd+bwd+SGD GB/s is reported beside it.

`value`    : tables, indices and cotangent already resident in HBM.
`e2e`      : the same step through the public host API with HOST buffers: indices and cotangent
             are copied in from pinned memory and the feature matrix is copied out, every step.
`roofline` : the dominant kernel's algorithmic bytes / its CUDA-event duration, vs the measured
             HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` / --impl reference: the CPU oracle (a C port of the reference's algorithm --
             the reference is Julia, which this image cannot run) on the box's host cores with
             the reference's threaded strategies, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "embeddingtables.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOAD = "C2: 26 tables x 1M rows x dim 128 f32, bag 32, batch 16384, PreallocationStrategy(prependrows=128): fwd + pullback + ensemble update!(Descent)"
NT, NROWS, DIM, BAG, BATCH, PREPEND = 26, 1_000_000, 128, 32, 16384, 128
ETA = 0.01
SEED = 0xE7AB1E + 2


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def make_indices(rng, dist, nt, nrows, bag, batch):
    """(bag, batch, nt) 1-based int64 indices, Julia 3-d container form."""
    if dist == "uniform":
        return rng.integers(1, nrows + 1, (bag, batch, nt), dtype=np.int64)
    # Zipf(alpha=1.05) on 1..nrows by inverse CDF; rank -> row through a seeded permutation per table
    w = 1.0 / np.arange(1, nrows + 1, dtype=np.float64) ** 1.05
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    out = np.empty((bag, batch, nt), dtype=np.int64, order="F")
    for t in range(nt):
        ranks = np.searchsorted(cdf, rng.random(bag * batch))
        perm = rng.permutation(nrows)
        out[:, :, t] = (perm[ranks] + 1).reshape(bag, batch, order="F")
    return out


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc = device, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(name)
        if sm:
            self.result = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                           "samples": len(sm)}


# ----------------------------------------------------------------------------------- CPU arm
def cpu_reference_arm(steps, warmup, dist, sample_tables=4):
    """The reference's CPU path (C port = oracle/) on this box's host cores: PreallocationStrategy
    forward (8 batch chunks x tables behind an atomic counter) + ensemble update! (index! per table,
    then 4 bucket splits x tables behind an atomic counter), all host threads.  Bounded sample:
    `sample_tables` of the 26 tables at the full batch/bag/dim (cost is linear in tables)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(SEED)
    nt = sample_tables
    tables = []
    for _ in range(nt):
        a = np.empty((DIM, NROWS), np.float32, order="F")
        a.reshape(-1, order="F")[:] = rng.random(DIM * NROWS, dtype=np.float32)
        tables.append(O.Table(a, static=True))
    I = make_indices(rng, dist, nt, NROWS, BAG, BATCH)
    Is = [np.asfortranarray(I[:, :, t]) for t in range(nt)]
    out = np.zeros((PREPEND + nt * DIM, BATCH), np.float32, order="F")
    delta = np.asfortranarray(rng.standard_normal(out.shape, dtype=np.float32))
    deltas = [delta[PREPEND + DIM * k: PREPEND + DIM * (k + 1)] for k in range(nt)]
    scratch = O.alloc_indexers(Is)  # caller-owned Indexers, reused like the reference's
    lookups = nt * BATCH * BAG
    t_fwd, t_upd = [], []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        O.maplookup("preallocation", tables, Is, prependrows=PREPEND, nthreads=cores, out=out)
        t1 = time.perf_counter()
        O.update_ensemble(tables, deltas, Is, ETA, num_splits=4, nthreads=cores, scratch=scratch)
        t2 = time.perf_counter()
        if s >= warmup:
            t_fwd.append(t1 - t0); t_upd.append(t2 - t1)
    step_s = float(np.mean(t_fwd) + np.mean(t_upd))
    return {"value": lookups / step_s, "unit": "lookups/s", "cores": cores, "kind": "port",
            "sample": f"{nt} of {NT} tables (1M x 128 f32), full batch {BATCH}, bag {BAG}, {dist} indices; "
                      f"{steps} timed steps after {warmup} warm-up; C port of the reference (Julia absent), "
                      f"AVX-512={bool(O.lib().etbo_uses_avx512())}",
            "ms_per_step": step_s * 1e3, "fwd_ms": float(np.mean(t_fwd)) * 1e3, "update_ms": float(np.mean(t_upd)) * 1e3,
            "lookups_per_step": lookups}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    r = cpu_reference_arm(steps, warmup, args.dist)
    line = {"impl": "reference", "metric": "embedding_lookups_per_sec", "value": r["value"], "unit": "lookups/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "dist": args.dist, "note": "bounded sample; lookups/s is per-table linear"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "fwd_ms": r["fwd_ms"], "update_ms": r["update_ms"], "gpu_launches": 0}
    emit_line(line)


# ----------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import embtab as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    E._lib.check(E.lib().etb_init(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        return run_sharded(args, rank, world, local)

    lib = E.lib()
    rng = np.random.default_rng(SEED)
    # ---- synthetic state, resident in HBM ------------------------------------------------
    gen = torch.Generator(device="cuda").manual_seed(SEED)
    tables = []
    for _ in range(NT):
        buf = torch.rand(DIM * NROWS, device="cuda", dtype=torch.float32, generator=gen)
        tables.append(E.SimpleEmbedding(E.DeviceArray(buf, (DIM, NROWS)), E.Static(DIM)))
    I_host = make_indices(rng, args.dist, NT, NROWS, BAG, BATCH)
    distinct = [int(np.unique(I_host[:, :, t]).size) for t in range(NT)]
    idx_pinned = E.pinned_empty((BAG, BATCH, NT), np.int64)
    idx_pinned[...] = I_host
    total_rows = PREPEND + NT * DIM
    out_pinned = E.pinned_empty((total_rows, BATCH), np.float32)
    delta_pinned = E.pinned_empty((total_rows, BATCH), np.float32)
    delta_pinned.reshape(-1, order="F")[:] = rng.standard_normal(total_rows * BATCH, dtype=np.float32)

    I_dev = E.DeviceArray.empty((BAG, BATCH, NT), np.int64).upload(idx_pinned)
    out_dev = E.DeviceArray.empty((total_rows, BATCH), np.float32)
    delta_dev = E.DeviceArray.empty((total_rows, BATCH), np.float32).upload(delta_pinned)
    strategy = E.PreallocationStrategy(PREPEND)
    indexer = E.Indexer()
    opt = E.Descent(ETA)
    S = E.Static(DIM)
    Is = list(E.colwrap(I_dev))
    launches = {"fwd": 0, "index": 0, "update": 0}

    def step(events=None, overlap=True):
        # index! needs only the indices: start it on a side stream so it overlaps the forward pass
        if overlap:
            E.prefetch_index(indexer, tables, Is)
        # forward: one fused launch writing straight into the concatenated matrix
        E.maplookup_(strategy, out_dev, tables, I_dev)
        launches["fwd"] = lib.etb_last_launch_count()
        if events: events[1].record()
        # backward: lazy pullback = row-slice views of the cotangent (no kernel)
        slicer = E.Slicer(PREPEND + 1, 1, delta_dev)
        grads = [E.SparseEmbeddingUpdate(S, slicer(DIM), i) for i in Is]
        # update!: index! (batched sort + bucket records) then fused segment-reduce + SGD
        if overlap:
            torch.cuda.current_stream().wait_event(indexer._event)
        else:
            E.index_(indexer, tables, grads)
            launches["index"] = lib.etb_last_launch_count()
        if events: events[2].record()
        E.update_(opt, tables, grads, [indexer]) if overlap else E.sparseupdate._apply(tables, grads, indexer, opt.eta)
        launches["update"] = lib.etb_last_launch_count()
        if events: events[3].record()

    # e2e: every step copies ITS indices and cotangent in from pinned host memory and its feature
    # matrix out.  Across steps the only pipelining is the standard input double buffer: step k+1's
    # indices are uploaded on a copy stream while step k's result is downloaded (H2D and D2H use
    # different copy engines).  Within a step the dependency chain is kept -- forward -> result on the
    # host -> cotangent from the host -> update -- and pipelined at its two PCIe legs:
    #   * the forward runs in E2E_CHUNKS column chunks, chunk c going to the host while chunk c+1 is looked up;
    #   * the cotangent comes in as E2E_GROUPS row slices (one group of tables each, a strided 2-D copy), and
    #     each group's update! (its own Indexer, index! prefetched beside the forward) runs as soon as its
    #     slice has landed, while the next slice is still on the wire.
    copy_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    E2E_CHUNKS = 4
    E2E_GROUPS = max(1, min(NT, int(os.environ.get("ETB_E2E_GROUPS", "13"))))
    bounds = [round(g * NT / E2E_GROUPS) for g in range(E2E_GROUPS + 1)]
    groups = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    group_ix = [E.Indexer() for _ in groups]
    I_buf = [I_dev, E.DeviceArray.empty((BAG, BATCH, NT), np.int64)]
    Is_buf = [Is, list(E.colwrap(I_buf[1]))]
    idx_ready = [None, None]
    buf_free = [None, None]
    e2e_state = {"k": 0}

    def upload_indices(slot):
        with torch.cuda.stream(copy_stream):
            if buf_free[slot] is not None:
                copy_stream.wait_event(buf_free[slot])    # the step that last used this buffer is done
            I_buf[slot].upload(idx_pinned)                # H2D: that step's indices
            idx_ready[slot] = copy_stream.record_event()

    def step_e2e():
        k = e2e_state["k"]
        slot = k % 2
        if idx_ready[slot] is None:
            upload_indices(slot)
        main = torch.cuda.current_stream()
        main.wait_event(idx_ready[slot])
        idx_ready[slot] = None
        for (a, b), ix in zip(groups, group_ix):          # side stream: overlaps forward + PCIe copies
            E.prefetch_index(ix, tables[a:b], Is_buf[slot][a:b])
        # forward in E2E_CHUNKS column chunks: chunk c's result goes to the host (D2H stream) while chunk
        # c+1 is being looked up
        for c in range(E2E_CHUNKS):
            c0, c1 = c * BATCH // E2E_CHUNKS, (c + 1) * BATCH // E2E_CHUNKS
            E.maplookup_(strategy, out_dev.cols(c0, c1), tables, [i.cols(c0, c1) for i in Is_buf[slot]])
            done = main.record_event()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                out_dev.cols(c0, c1).download(out_pinned[:, c0:c1])   # D2H: the step's result, chunk c
        upload_indices(1 - slot)                          # next step's indices, beside this step's D2H
        # H2D: the upstream cotangent, once the host has the whole feature matrix; group g's rows
        # (the first slice also carries the PREPEND rows of the dense part: the whole matrix is copied)
        landed = []
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_stream(d2h_stream)
            for a, b in groups:
                r0, r1 = (0 if a == 0 else PREPEND + a * DIM), PREPEND + b * DIM
                delta_dev.rows(r0, r1).upload(delta_pinned[r0:r1])
                landed.append(copy_stream.record_event())
        slicer = E.Slicer(PREPEND + 1, 1, delta_dev)
        grads = [E.SparseEmbeddingUpdate(S, slicer(DIM), i) for i in Is_buf[slot]]
        for (a, b), ix, ev in zip(groups, group_ix, landed):
            main.wait_event(ev)
            E.update_(opt, tables[a:b], grads[a:b], [ix])
        buf_free[slot] = main.record_event()
        e2e_state["k"] = k + 1

    def sync():
        torch.cuda.synchronize()

    overlap = not args.no_overlap
    # nvidia-smi needs up to a second to deliver its first sample: start it before the warm-up so that it is
    # sampling (every 20 ms) throughout the timed regions; stopped after the e2e region
    clocks = ClockSampler(local)
    clocks.__enter__()
    for _ in range(args.warmup):
        step(overlap=False)
        step(overlap=overlap)
    sync()
    K = args.steps
    # ---- the timed region: K steps, index! overlapped with the forward unless --no-overlap
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    start.record()
    for k in range(K):
        step(overlap=overlap)
    end.record()
    sync()
    ms_per_step = start.elapsed_time(end) / K
    # ---- per-kernel times: the same step with the phases back to back (CUDA events on the stream)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    for k in range(K):
        ev[k][0].record()
        step(ev[k], overlap=False)
    sync()
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    index_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    upd_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
    serial_ms = float(np.mean([e[0].elapsed_time(e[3]) for e in ev]))
    lookups = NT * BATCH * BAG

    # ---- e2e: host buffers in, host result out, every step --------------------------------
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        step_e2e()
    e1.record()
    sync()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3 / K
    e2e_ms = max(e0.elapsed_time(e1) / K, e2e_wall_ms)  # host-side work counts too
    clocks.__exit__(None, None, None)

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    s_, b_ = 4, 8
    fwd_bytes = NT * BATCH * (BAG * (b_ + DIM * s_) + DIM * s_)                  # pooled fwd
    u_sum = sum(distinct)
    upd_kernel_bytes = NT * BATCH * DIM * s_ + 2 * u_sum * DIM * s_ + NT * BATCH * BAG * 4   # delta + row RMW + map
    upd_total_bytes = NT * BATCH * BAG * b_ + NT * BATCH * DIM * s_ + 2 * u_sum * DIM * s_  # update
    kernels = {
        "pooled_kernel": {"ms": fwd_ms, "bytes": fwd_bytes, "gbs": fwd_bytes / fwd_ms / 1e6},
        "sgd_update_kernel": {"ms": upd_ms, "bytes": upd_kernel_bytes, "gbs": upd_kernel_bytes / upd_ms / 1e6},
        "index(make_pairs+radix sort+select)": {"ms": index_ms},
    }
    dom = "pooled_kernel" if fwd_ms >= upd_ms else "sgd_update_kernel"
    traffic = None      # DRAM bytes per launch of that kernel from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_kernels.json")) as f:
            traffic = json.load(f)["kernels"][dom]["dram_bytes"] if args.dist == "uniform" else None
    except Exception:
        traffic = None
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
            "frac_of_nominal_8TBs": kernels[dom]["gbs"] / 8000.0}

    cpu = cpu_reference_arm(steps=2, warmup=1, dist=args.dist) if not args.no_cpu_baseline else None
    gpu_launches = (launches["fwd"] + launches["index"] + launches["update"]) * K
    line = {
        "metric": "embedding_lookups_per_sec", "value": lookups / (ms_per_step * 1e-3), "unit": "lookups/s",
        "n_gpus": 1, "steps": K, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "dist": args.dist, "index_type": "int64", "eta": ETA,
                   "schedule": "index! on a side stream overlapping the forward" if overlap else "phases back to back",
                   "ms_per_step_phases_back_to_back": serial_ms,
                   "l2": "inputs larger than L2: 13.3 GB of tables, random rows; no flush needed",
                   "distinct_rows_per_table_mean": u_sum / NT},
        "e2e": {"value": lookups / (e2e_ms * 1e-3), "unit": "lookups/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(idx_pinned.nbytes + delta_pinned.nbytes),
                "d2h_bytes_per_step": int(out_pinned.nbytes),
                "pipeline": "indices double-buffered; forward in %d column chunks with overlapped D2H; cotangent in %d "
                            "table-group row slices (2-D H2D), each group's update! as its slice lands"
                            % (E2E_CHUNKS, len(groups))},
        "gpu_launches": gpu_launches,
        "launches_per_step": launches,
        "clocks": clocks.result,
        "roofline": roof,
        "kernels": kernels,
        "fwd_lookups_per_sec": lookups / (fwd_ms * 1e-3),
        "fwd_bwd_sgd_gbs": (fwd_bytes + upd_total_bytes) / (ms_per_step * 1e6),
        "fwd_bwd_sgd_frac_of_peak": (fwd_bytes + upd_total_bytes) / (ms_per_step * 1e6) / peak,
        "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
    }
    emit_line(line)


def run_sharded(args, rank, world, local):
    """N > 1: table-wise sharding (north star / SURVEY 8e).  Weak scaling: every rank owns 26 tables
    (26*N in the ensemble), the global batch stays 16384, every rank looks its tables up for the
    whole global batch, the all-to-all gives each rank all 26*N tables for its 16384/N samples, the
    reverse all-to-all carries the cotangent back and each owner updates its tables.  Per-GPU
    lookup and update work equal the N=1 run's; value = all ranks' lookups / max-over-ranks time."""
    import torch
    import torch.distributed as dist

    import embtab as E
    from embtab.dist import ShardedEnsemble, ShardPlan

    lib = E.lib()
    rng = np.random.default_rng(SEED + 1000 * rank)
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    tables = []
    for _ in range(NT):
        buf = torch.rand(DIM * NROWS, device="cuda", dtype=torch.float32, generator=gen)
        tables.append(E.SimpleEmbedding(E.DeviceArray(buf, (DIM, NROWS)), E.Static(DIM)))
    plan = ShardPlan([DIM] * (NT * world), world, rank, PREPEND, BATCH)
    ens = ShardedEnsemble(tables, plan, fused=not args.nccl_a2a)
    I_host = make_indices(rng, args.dist, NT, NROWS, BAG, BATCH)
    idx_pinned = E.pinned_empty((BAG, BATCH, NT), np.int64)
    idx_pinned[...] = I_host
    I_dev = E.DeviceArray.empty((BAG, BATCH, NT), np.int64).upload(idx_pinned)
    out_shape = (plan.total_rows, plan.my_cols)
    out_pinned = E.pinned_empty(out_shape, np.float32)
    delta_pinned = E.pinned_empty(out_shape, np.float32)
    delta_pinned.reshape(-1, order="F")[:] = rng.standard_normal(out_shape[0] * out_shape[1], dtype=np.float32)
    delta_dev = E.DeviceArray.empty(out_shape, np.float32).upload(delta_pinned)
    opt = E.Descent(ETA)
    launches = [0]

    def step(events=None):
        ens.forward(I_dev)
        n = ens.launches
        if events: events[1].record()
        grads = ens.backward(delta_dev)
        n += 1
        if events: events[2].record()
        ens.update_(opt, grads)
        n += lib.etb_last_launch_count() + ens.index_launches   # update kernels + the prefetched index! launches
        if events: events[3].record()
        launches[0] = n

    # e2e: indices and cotangent in from pinned host memory, this rank's feature-matrix columns out, every step.
    # Not pipelined: with N ranks pulling through one host the PCIe legs are host-bound (double-buffering the
    # index upload was measured at N = 2 and changed nothing).
    def step_e2e():
        I_dev.upload(idx_pinned)
        ens.forward(I_dev)
        ens.out.download(out_pinned)
        delta_dev.upload(delta_pinned)
        ens.update_(opt, ens.backward(delta_dev))

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)   # started before the warm-up (nvidia-smi needs ~1 s for its first sample),
    clocks.__enter__()             # stopped after the e2e region
    for _ in range(args.warmup):
        step()
    sync()
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    end = torch.cuda.Event(enable_timing=True)
    sync()
    for k in range(K):
        ev[k][0].record()
        step(ev[k])
    end.record()
    sync()
    t = torch.tensor([ev[0][0].elapsed_time(end) / K,
                      float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
                      float(np.mean([e[1].elapsed_time(e[2]) for e in ev])),
                      float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # max over ranks, device-timed
    ms_per_step, fwd_ms, bwd_ms, upd_ms = t.tolist()

    for _ in range(2):
        step_e2e()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        step_e2e()
    e1.record()
    sync()
    e2e = torch.tensor([max(e0.elapsed_time(e1) / K, (time.perf_counter() - t0) * 1e3 / K)], device="cuda", dtype=torch.float64)
    clocks.__exit__(None, None, None)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_ms = e2e.item()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        lookups = world * NT * BATCH * BAG
        s_, b_ = 4, 8
        fwd_bytes = NT * BATCH * (BAG * (b_ + DIM * s_) + DIM * s_)
        a2a_bytes = plan.my_rows * BATCH * s_ * (world - 1) / world          # sent per rank per direction
        line = {
            "metric": "embedding_lookups_per_sec", "value": lookups / (ms_per_step * 1e-3), "unit": "lookups/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + f"; table-wise sharded: {NT} tables per GPU ({NT * world} total), global batch "
                       f"{BATCH}, exchange = " + ("NCCL all-to-all + pack/unpack" if args.nccl_a2a else
                       "fused: lookup/scatter kernels store into peer HBM over NVLink (CUDA IPC), all-reduce barrier"),
                       "dist": args.dist, "index_type": "int64",
                       "eta": ETA, "l2": "inputs larger than L2: 13.3 GB of tables per GPU, random rows; no flush needed"},
            "e2e": {"value": lookups / (e2e_ms * 1e-3), "unit": "lookups/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(idx_pinned.nbytes + delta_pinned.nbytes),
                    "d2h_bytes_per_step": int(out_pinned.nbytes)},
            "gpu_launches": launches[0] * K, "clocks": clocks.result,
            "roofline": {"bound": "hbm", "kernel": "pooled_kernel+a2a (fwd phase, max over ranks)",
                         "achieved": fwd_bytes / fwd_ms / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": fwd_bytes / fwd_ms / 1e6 / peak, "traffic": None, "peak_source": peak_src},
            "phases_ms": {"fwd_lookup+exchange": fwd_ms, "bwd_exchange": bwd_ms, "index+update": upd_ms},
            "nvlink": {"bytes_sent_per_rank_per_direction": a2a_bytes, "peak_gbs": 770.0,
                       "note": "phase times include the lookup / pack kernels; see profiles/ for the split"},
            "cpu_baseline": None,
        }
        emit_line(line)
    dist.destroy_process_group()


_JSON_OUT = None


def emit_line(line):
    """the ONE JSON line of the contract, on the process's real stdout"""
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def main():
    # Libraries print to stdout behind our back (NCCL's "NCCL version ..." banner under NCCL_DEBUG=VERSION/WARN):
    # keep the real stdout for the JSON line and point file descriptor 1 at stderr for everything else.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dist", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run index! after the forward instead of beside it")
    ap.add_argument("--nccl-a2a", action="store_true",
                    help="N>1: exchange with NCCL all-to-all + pack/unpack instead of fused NVLink peer stores")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
