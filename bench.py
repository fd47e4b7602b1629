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
    """nvidia-smi clocks/throttle reasons DURING the timed regions (B200_PROFILING.md recipe).  nvidia-smi needs seconds
    for its first sample on an 8-GPU box, so it is started when the process starts (start()); every sample carries
    nvidia-smi's own timestamp and only those between mark_begin() and stop() -- warm-up, timed region, per-phase pass
    and e2e region -- are reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.t_begin = device, None, None
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def mark_begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()

    def __enter__(self):          # start + mark in one go (single-process callers that create the sampler late)
        if self.proc is None:
            self.start()
        self.mark_begin()
        return self

    def __exit__(self, *a):
        self.stop()

    def stop(self):
        import datetime
        if self.proc is None:
            return
        time.sleep(0.15)
        t_end = datetime.datetime.now()
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return
        self.proc = None
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(f[2]), float(f[3]), {n for n, v in zip(names, f[6:10]) if v.lower() == "active"}))
            except ValueError:
                continue
        lo = (self.t_begin or t_end) - datetime.timedelta(milliseconds=50)
        window = [r for r in rows if lo <= r[0] <= t_end]
        where = "timed regions"
        if not window and rows:   # clock skew between nvidia-smi's stamps and ours, or a sampling gap: the newest samples
            window, where = rows[-10:], "last samples before the end of the timed regions"
        if window:
            reasons = set().union(*[r[3] for r in window])
            self.result = {"sm_mhz": float(np.median([r[1] for r in window])), "sm_max_mhz": float(max(r[2] for r in window)),
                           "reasons": sorted(reasons), "samples": len(window), "window": where}


def common_config(world, dist):
    """`config` of the JSON line: what the workload IS.  Both arms (ours / --impl reference) emit exactly this
    dict for the same --gpus/--dist; how each arm schedules it is reported outside `config`."""
    nt = NT * world
    w = WORKLOAD if world == 1 else (
        WORKLOAD + f"; weak scaling over {world} GPUs: {NT} tables per GPU ({nt} in the ensemble), global batch {BATCH}, "
        "tables block-partitioned table-wise, pooled outputs / cotangents exchanged all-to-all")
    return {"workload": w, "dist": dist, "tables": nt, "rows": NROWS, "dim": DIM, "bag": BAG, "batch": BATCH,
            "prependrows": PREPEND, "index_type": "int64", "eta": ETA,
            "l2": "inputs larger than L2: 13.3 GB of tables per GPU, random rows; no flush needed"}


# ----------------------------------------------------------------------------------- CPU arm
def cpu_reference_arm(steps, warmup, dist, world=1):
    """The reference's CPU path (C port = oracle/; the reference is Julia, which this image cannot run) on this
    box's host cores, on the FULL ensemble of the workload: PreallocationStrategy forward (8 batch chunks x tables
    behind an atomic counter, reference src/lookup.jl:316-371) + ensemble update! (index! per table in parallel,
    then num_splits = 4 bucket splits x tables behind an atomic counter, src/sparseupdate.jl:199-238), one pinned
    worker thread per core.  Every table is there because the reference's index! phase is table-parallel only
    (:211-213): a sample of a few tables would leave most cores idle in the phase that dominates the step.
    The update is timed with both Indexer flavours of the reference (SparseIndexer = its default, DenseIndexer);
    the faster one is reported."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    cores = O.allowed_cpus()
    O.set_pinning(True)
    nt_full = NT * world
    nt = nt_full
    try:   # N > 1 (weak scaling: 26 tables per GPU): as many of the tables as half of the free host RAM holds
        import psutil
        per_table = DIM * NROWS * 4 + 5 * BAG * BATCH * 8 + 2 * DIM * BATCH * 4
        nt = int(max(NT, min(nt_full, psutil.virtual_memory().available * 0.5 // per_table)))
    except Exception:
        nt = min(nt_full, NT)
    rng = np.random.default_rng(SEED)
    tables = []
    for t in range(nt):
        a = np.empty((DIM, NROWS), np.float32, order="F")
        O.fill_uniform(a, SEED + t, cores)          # uniform [0, 1) like the reference's tests; first touch by the workers
        tables.append(O.Table(a, static=True))
    Is = []
    for t0 in range(0, nt, NT):                      # same generator as the GPU arm, NT tables at a time
        I = make_indices(rng, dist, min(NT, nt - t0), NROWS, BAG, BATCH)
        Is += [np.asfortranarray(I[:, :, t]) for t in range(I.shape[2])]
    out = np.zeros((PREPEND + nt * DIM, BATCH), np.float32, order="F")
    delta = np.empty(out.shape, np.float32, order="F")
    O.fill_uniform(delta, SEED - 1, cores)           # values do not matter for the timing
    delta -= 0.5
    deltas = [delta[PREPEND + DIM * k: PREPEND + DIM * (k + 1)] for k in range(nt)]
    scratch = O.alloc_indexers(Is)                   # caller-owned Indexers, reused like the reference's
    lookups = nt * BATCH * BAG
    t_fwd, t_upd = [], {False: [], True: []}
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        O.maplookup("preallocation", tables, Is, prependrows=PREPEND, nthreads=cores, out=out)
        t1 = time.perf_counter()
        O.update_ensemble(tables, deltas, Is, ETA, num_splits=4, nthreads=cores, scratch=scratch, dense=False)
        t2 = time.perf_counter()
        if s >= warmup:
            t_fwd.append(t1 - t0); t_upd[False].append(t2 - t1)
    for s in range(1 + min(steps, 3)):               # the DenseIndexer flavour of the same update
        t1 = time.perf_counter()
        O.update_ensemble(tables, deltas, Is, ETA, num_splits=4, nthreads=cores, scratch=scratch, dense=True)
        if s >= 1:
            t_upd[True].append(time.perf_counter() - t1)
    upd = {k: float(np.mean(v)) for k, v in t_upd.items()}
    dense = upd[True] < upd[False]
    step_s = float(np.mean(t_fwd)) + upd[dense]
    sample = (f"all {nt} tables" if nt == nt_full else f"{nt} of {nt_full} tables (host RAM)") + \
        (f" (1M x 128 f32, {nt * DIM * NROWS * 4 / 1e9:.1f} GB), full batch {BATCH}, bag {BAG}, {dist} indices; "
         f"{steps} timed steps after {warmup} warm-up; {cores} pinned threads; C port of the reference (Julia absent), "
         f"AVX-512={bool(O.lib().etbo_uses_avx512())}; update! with the {'Dense' if dense else 'Sparse'}Indexer "
         f"(sparse {upd[False] * 1e3:.1f} ms, dense {upd[True] * 1e3:.1f} ms)")
    return {"value": lookups / step_s, "unit": "lookups/s", "cores": cores, "kind": "port", "sample": sample,
            "ms_per_step": step_s * 1e3, "fwd_ms": float(np.mean(t_fwd)) * 1e3, "update_ms": upd[dense] * 1e3,
            "lookups_per_step": lookups, "tables": nt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_reference_arm(steps, warmup, args.dist, world)
    line = {"impl": "reference", "metric": "embedding_lookups_per_sec", "value": r["value"], "unit": "lookups/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_config(world, args.dist),
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
    args._clocks = ClockSampler(local).start()   # sampling from now on; the report keeps the timed regions' samples
    args._numa = bind_to_gpu_numa(local) if (world > 1 or os.environ.get("ETB_BIND_NUMA")) else None
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
    # indices are uploaded on the H2D stream while step k is in flight.  Within a step the dependency
    # chain is kept PER SAMPLE (= per column of the feature matrix; what a DLRM's per-sample loss gives):
    # forward of column j -> column j of the result on the host -> column j of the cotangent from the
    # host -> update! of a table once ALL columns of its cotangent rows are in HBM.  PCIe is full duplex
    # (tools/pcie_probe.py: 55 GB/s one way, 2 x 46 GB/s both ways at once), so the schedule keeps both
    # directions busy:
    #   * the forward runs in E2E_CHUNKS column chunks, chunk c going to the host while chunk c+1 is looked up;
    #   * the cotangent of the first E2E_DENSE column chunks comes in as whole column chunks (contiguous),
    #     each as soon as ITS result chunk has reached the host -- beside the later result chunks' D2H;
    #   * the remaining columns come in as E2E_GROUPS row slices (one group of tables each, a strided 2-D
    #     copy), and each group's update! (its own Indexer, index! prefetched beside the forward) runs as
    #     soon as its slice has landed, while the next slice is still on the wire.
    # ETB_E2E_DUPLEX=0 is round 1's schedule: the whole result on the host before any cotangent is sent.
    copy_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    DUPLEX = os.environ.get("ETB_E2E_DUPLEX", "1") != "0"
    E2E_CHUNKS = max(1, int(os.environ.get("ETB_E2E_CHUNKS", "8" if DUPLEX else "4")))
    E2E_DENSE = min(E2E_CHUNKS - 1, max(0, int(os.environ.get("ETB_E2E_DENSE", str(E2E_CHUNKS // 2))))) if DUPLEX else 0
    E2E_GROUPS = max(1, min(NT, int(os.environ.get("ETB_E2E_GROUPS", "7"))))
    bounds = [round(g * NT / E2E_GROUPS) for g in range(E2E_GROUPS + 1)]
    groups = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    group_ix = [E.Indexer() for _ in groups]
    cb = [c * BATCH // E2E_CHUNKS for c in range(E2E_CHUNKS + 1)]
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
        if DUPLEX:
            upload_indices(1 - slot)                      # next step's indices: first in the H2D queue, beside the forward
        # forward in E2E_CHUNKS column chunks: chunk c's result goes to the host (D2H stream) while chunk
        # c+1 is being looked up; the cotangent of a dense chunk follows its result chunk (H2D stream)
        for c in range(E2E_CHUNKS):
            c0, c1 = cb[c], cb[c + 1]
            E.maplookup_(strategy, out_dev.cols(c0, c1), tables, [i.cols(c0, c1) for i in Is_buf[slot]])
            done = main.record_event()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                out_dev.cols(c0, c1).download(out_pinned[:, c0:c1])   # D2H: the step's result, chunk c
                on_host = d2h_stream.record_event()
            if c < E2E_DENSE:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(on_host)       # these samples' results are on the host
                    delta_dev.cols(c0, c1).upload(delta_pinned[:, c0:c1])
        if not DUPLEX:
            upload_indices(1 - slot)                      # next step's indices, beside this step's D2H
        # H2D: the rest of the upstream cotangent, once the host has the whole feature matrix; group g's rows
        # (the first slice also carries the PREPEND rows of the dense part: the whole matrix is copied)
        landed = []
        t0 = cb[E2E_DENSE]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_stream(d2h_stream)
            for a, b in groups:
                r0, r1 = (0 if a == 0 else PREPEND + a * DIM), PREPEND + b * DIM
                delta_dev.rows(r0, r1).cols(t0, BATCH).upload(delta_pinned[r0:r1, t0:])
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
    clocks = args._clocks
    clocks.mark_begin()
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
    clocks.stop()

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
        "index(segmented radix sort: 2 x (hist, scan, scatter) + bucket records)": {"ms": index_ms},
    }
    dom = "pooled_kernel" if fwd_ms >= upd_ms else "sgd_update_kernel"
    traffic = None      # DRAM bytes per launch of that kernel from the newest committed ncu --set full capture
    try:                # (profiles/r2_ncu_kernels.json, regenerated by tools/final_n1.sh whenever a kernel changes)
        with open(os.path.join(ROOT, "profiles", "r2_ncu_kernels.json")) as f:
            key = {"sgd_update_kernel": "sgd_update_exact_kernel", "pooled_kernel": "pooled_kernel"}[dom]
            traffic = json.load(f)["kernels"][key]["dram_bytes"] if args.dist == "uniform" else None
    except Exception:
        traffic = None
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
            "frac_of_nominal_8TBs": kernels[dom]["gbs"] / 8000.0}

    cpu = cpu_reference_arm(steps=3, warmup=1, dist=args.dist) if not args.no_cpu_baseline else None
    gpu_launches = (launches["fwd"] + launches["index"] + launches["update"]) * K
    line = {
        "metric": "embedding_lookups_per_sec", "value": lookups / (ms_per_step * 1e-3), "unit": "lookups/s",
        "n_gpus": 1, "steps": K, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(1, args.dist),
        "schedule": "index! on a side stream overlapping the forward" if overlap else "phases back to back",
        "ms_per_step_phases_back_to_back": serial_ms,
        "distinct_rows_per_table_mean": u_sum / NT,
        "e2e": {"value": lookups / (e2e_ms * 1e-3), "unit": "lookups/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(idx_pinned.nbytes + delta_pinned.nbytes),
                "d2h_bytes_per_step": int(out_pinned.nbytes),
                "pipeline": ("indices double-buffered; forward in %d column chunks with overlapped D2H; cotangent: the first %d "
                             "column chunks each right behind its own result chunk (PCIe full duplex, per-sample dependency), the "
                             "remaining columns in %d table-group row slices (2-D H2D), each group's update! as its slice lands"
                             % (E2E_CHUNKS, E2E_DENSE, len(groups))) if DUPLEX else
                            ("indices double-buffered; forward in %d column chunks with overlapped D2H; whole result on the host, then "
                             "the cotangent in %d table-group row slices (2-D H2D), each group's update! as its slice lands"
                             % (E2E_CHUNKS, len(groups)))},
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


def bind_to_gpu_numa(local):
    """Pin this process (and so the first touch of its pinned buffers) to the CPUs of the NUMA node the GPU hangs
    off: with N ranks pulling host buffers through one box, remote-node traffic is what saturates first."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # no sysfs / no such attribute: run unbound
        return {"numa_node": None, "error": str(e)[:80]}


def rank_delta_block(q, owner, rows, cols):
    """cotangent rows of `owner`'s tables at rank q (rows x cols of q's batch slice): a seed per (rank, owner) so
    that any rank can regenerate what a peer sent it (the self-check below)"""
    return np.random.default_rng(SEED + 7919 * (q + 1) + owner).standard_normal((rows, cols), dtype=np.float32)


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
    numa = getattr(args, "_numa", None)
    rng = np.random.default_rng(SEED + 1000 * rank)
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    tables = []
    for _ in range(NT):
        buf = torch.rand(DIM * NROWS, device="cuda", dtype=torch.float32, generator=gen)
        tables.append(E.SimpleEmbedding(E.DeviceArray(buf, (DIM, NROWS)), E.Static(DIM)))
    plan = ShardPlan([DIM] * (NT * world), world, rank, PREPEND, BATCH)
    fused = not args.nccl_a2a
    G = max(1, int(os.environ.get("ETB_TABLE_GROUPS", "1"))) if fused else 1   # measured on 8 GPUs: 1 group 3.64 ms, 2 groups 3.67, 4 groups 4.47
    # the host-buffer (e2e) step always runs by table groups (each group's exchange + update! as its cotangent slice
    # lands); the device-timed step runs the ensemble whole unless ETB_TABLE_GROUPS asks for groups
    GE = (G if G > 1 else max(1, int(os.environ.get("ETB_E2E_TABLE_GROUPS", "4")))) if fused else 1
    ens = ShardedEnsemble(tables, plan, fused=fused, table_groups=GE, peer_barrier=not args.nccl_barrier,
                          copy_engine=args.exchange == "copy")
    GE = ens.n_groups
    DEVG = G > 1                      # device-timed step by groups?
    G = GE if DEVG else 1
    I_host = make_indices(rng, args.dist, NT, NROWS, BAG, BATCH)
    idx_pinned = E.pinned_empty((BAG, BATCH, NT), np.int64)
    idx_pinned[...] = I_host
    I_dev = E.DeviceArray.empty((BAG, BATCH, NT), np.int64).upload(idx_pinned)
    out_shape = (plan.total_rows, plan.my_cols)
    out_pinned = E.pinned_empty(out_shape, np.float32)
    delta_pinned = E.pinned_empty(out_shape, np.float32)
    delta_pinned[:PREPEND] = 0.0
    for o in range(world):
        delta_pinned[plan.row_off[o]:plan.row_off[o] + plan.rows[o]] = rank_delta_block(rank, o, plan.rows[o], plan.my_cols)
    delta_dev = E.DeviceArray.empty(out_shape, np.float32).upload(delta_pinned)
    opt = E.Descent(ETA)
    launches = [0]

    # index! needs only the indices: it runs on a side stream beside the forward ("forward") or beside the backward
    # exchange ("exchange": that phase is NVLink-bound and leaves HBM idle, while lookup and index! compete for it)
    IDX_BESIDE = args.index_beside

    def step(events=None, pipelined=True):
        ens.forward(I_dev, prefetch_index=IDX_BESIDE == "forward", grouped=DEVG)
        n = ens.launches + (1 if ens.peer_barrier else 0)
        if events: events[1].record()
        if IDX_BESIDE == "exchange":
            ens.prefetch_index(DEVG)
        if pipelined:      # backward exchange and update!, table group by table group
            ens.backward_update_(opt, delta_dev, grouped=DEVG)
            n += ens.update_launches + ens.index_launches + 2 * G
            if events: events[2].record(); events[3].record()
        else:
            grads = ens.backward(delta_dev)
            n += 2
            if events: events[2].record()
            ens.update_(opt, grads, grouped=DEVG)
            n += lib.etb_last_launch_count() * G + ens.index_launches
            if events: events[3].record()
        launches[0] = n

    # ---- self-check before anything is timed (uniform runs): one table block that a PEER looked up for me and one of
    # MY tables after a whole step, against a single-GPU recomputation from regenerated inputs
    check = "skipped"
    if args.dist == "uniform" and not args.no_self_check:
        q = (rank + 1) % world                                            # a peer: its first table, my columns
        genq = torch.Generator(device="cuda").manual_seed(SEED + q)
        tq = E.SimpleEmbedding(E.DeviceArray(torch.rand(DIM * NROWS, device="cuda", dtype=torch.float32, generator=genq),
                                             (DIM, NROWS)), E.Static(DIM))
        Iq = make_indices(np.random.default_rng(SEED + 1000 * q), args.dist, NT, NROWS, BAG, BATCH)[:, plan.clo[rank]:plan.chi[rank], 0]
        t0_before = tables[0].to_numpy()
        step()
        torch.cuda.synchronize()
        want = E.lookup(tq, np.asfortranarray(Iq)).numpy()
        got = ens.out.rows(plan.row_off[q], plan.row_off[q] + DIM).numpy()
        ok = np.array_equal(got, want)
        del tq
        ref = E.SimpleEmbedding(t0_before, E.Static(DIM))                 # my first table, updated on one GPU
        dglob = np.concatenate([rank_delta_block(p_, rank, plan.rows[rank], plan.cols[p_])[:DIM] for p_ in range(world)], axis=1)
        E.update_(opt, ref, E.SparseEmbeddingUpdate(E.Static(DIM), np.asfortranarray(dglob), np.asfortranarray(I_host[:, :, 0])))
        torch.cuda.synchronize()
        ok = ok and np.array_equal(ref.to_numpy(), tables[0].to_numpy())
        flag = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flag)
        check = "ok" if flag.item() == 0 else "FAILED"
        print(f"[rank {rank}] dist check {'ok' if ok else 'FAILED'}: block of peer {q} and my updated table 0 "
              f"{'equal' if ok else 'DIFFER from'} the single-GPU recomputation", file=sys.stderr, flush=True)
        if check != "ok":
            raise SystemExit("bench.py: sharded self-check failed")
        del ref
        torch.cuda.empty_cache()

    # ---- e2e: indices and cotangent in from pinned host memory, this rank's feature-matrix columns out, every
    # step, pipelined like the N = 1 path: indices double-buffered (step k+1's upload beside step k's download), the
    # forward in column chunks whose D2H overlaps the next chunk's lookup, the cotangent in table-group row slices
    # (one strided copy per owner and group), each group's exchange + update! as its slice has landed.
    # With ETB_E2E_DUPLEX (default) the cotangent of the first E2E_DENSE column chunks follows each chunk's own result
    # (per-sample dependency, PCIe full duplex) exactly as at N = 1.
    copy_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    DUPLEX = fused and os.environ.get("ETB_E2E_DUPLEX", "1") != "0"
    E2E_CHUNKS = max(1, int(os.environ.get("ETB_E2E_CHUNKS", "8" if DUPLEX else "4"))) if fused else 1
    E2E_DENSE = min(E2E_CHUNKS - 1, max(0, int(os.environ.get("ETB_E2E_DENSE", str(E2E_CHUNKS // 2))))) if DUPLEX else 0
    I_buf = [I_dev, E.DeviceArray.empty((BAG, BATCH, NT), np.int64)]
    idx_ready, buf_free, e2e_state = [None, None], [None, None], {"k": 0}
    cb = [round(c * plan.my_cols / E2E_CHUNKS) for c in range(E2E_CHUNKS + 1)]

    def upload_indices(slot):
        with torch.cuda.stream(copy_stream):
            if buf_free[slot] is not None:
                copy_stream.wait_event(buf_free[slot])
            I_buf[slot].upload(idx_pinned)
            idx_ready[slot] = copy_stream.record_event()

    def step_e2e():
        if not fused:      # NCCL variant: the plain chain
            I_dev.upload(idx_pinned)
            ens.forward(I_dev)
            ens.out.download(out_pinned)
            delta_dev.upload(delta_pinned)
            ens.update_(opt, ens.backward(delta_dev))
            return
        k = e2e_state["k"]
        slot = k % 2
        if idx_ready[slot] is None:
            upload_indices(slot)
        main = torch.cuda.current_stream()
        main.wait_event(idx_ready[slot])
        idx_ready[slot] = None
        if DUPLEX:
            upload_indices(1 - slot)                       # next step's indices: first in the H2D queue
        for c in range(E2E_CHUNKS):
            ens.forward(I_buf[slot], cols=(cb[c], cb[c + 1]), prefetch_index=(c == 0), grouped=True)
            done = main.record_event()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                ens.out.cols(cb[c], cb[c + 1]).download(out_pinned[:, cb[c]:cb[c + 1]])
                on_host = d2h_stream.record_event()
            if c < E2E_DENSE and cb[c + 1] > cb[c]:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(on_host)        # these samples' results are on the host
                    delta_dev.cols(cb[c], cb[c + 1]).upload(delta_pinned[:, cb[c]:cb[c + 1]])
        if not DUPLEX:
            upload_indices(1 - slot)
        landed = []
        t0 = cb[E2E_DENSE]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_stream(d2h_stream)            # the host has the whole result before the rest of the cotangent exists
            for g in range(GE):
                for o in range(world):
                    r0 = plan.row_off[o] + ens.group_rows[o][g][0]
                    r1 = r0 + ens.group_rows[o][g][1]
                    if g == 0 and o == 0:
                        r0 = 0                             # the dense part's rows travel too: the whole matrix is copied
                    delta_dev.rows(r0, r1).cols(t0, plan.my_cols).upload(delta_pinned[r0:r1, t0:])
                landed.append(copy_stream.record_event())
        ens.update_launches = 0
        for g in range(GE):
            main.wait_event(landed[g])
            ens.scatter_group(delta_dev, g)
            ens.update_group_(opt, g)
        ens.join_updates()
        buf_free[slot] = main.record_event()
        e2e_state["k"] = k + 1

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    clocks = args._clocks          # sampling since the process started; stopped after the e2e region
    clocks.mark_begin()
    for _ in range(args.warmup):
        step()
    sync()
    K = args.steps
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    start.record()
    for k in range(K):
        step()
    end.record()
    sync()
    # phases: the same step with the backward exchange and the update back to back (not pipelined)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    for k in range(K):
        ev[k][0].record()
        step(ev[k], pipelined=False)
    sync()
    t = torch.tensor([start.elapsed_time(end) / K,
                      float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
                      float(np.mean([e[1].elapsed_time(e[2]) for e in ev])),
                      float(np.mean([e[2].elapsed_time(e[3]) for e in ev])),
                      float(np.mean([e[0].elapsed_time(e[3]) for e in ev]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # max over ranks, device-timed
    ms_per_step, fwd_ms, bwd_ms, upd_ms, serial_ms = t.tolist()

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
    clocks.stop()
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
            "config": common_config(world, args.dist),
            "exchange": ("NCCL all-to-all + pack/unpack kernels" if args.nccl_a2a else
                         ("fused: lookup / scatter kernels store into peer HBM over NVLink (CUDA IPC); " if not ens.copy_engine else
                          "fused: kernels write local staging blocks, copy engines push them into peer HBM over NVLink (CUDA IPC); ") +
                         ("peer-memory flag barrier" if ens.peer_barrier else "NCCL all-reduce barrier") +
                         f"; backward exchange + update! pipelined over {G} table groups; index! beside the {IDX_BESIDE}"),
            "self_check": check, "numa": numa,
            "e2e": {"value": lookups / (e2e_ms * 1e-3), "unit": "lookups/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(idx_pinned.nbytes + delta_pinned.nbytes),
                    "d2h_bytes_per_step": int(out_pinned.nbytes),
                    "pipeline": ("indices double-buffered; forward in %d column chunks with overlapped D2H; cotangent: the first %d column "
                                 "chunks each right behind its own result chunk (full duplex), the rest in %d table-group row slices, each "
                                 "group's exchange + update! as its slice lands" % (E2E_CHUNKS, E2E_DENSE, GE))
                                if fused else "plain chain"},
            "gpu_launches": launches[0] * K, "clocks": clocks.result,
            "roofline": {"bound": "hbm", "kernel": "pooled_kernel+a2a (fwd phase, max over ranks)",
                         "achieved": fwd_bytes / fwd_ms / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": fwd_bytes / fwd_ms / 1e6 / peak, "traffic": None, "peak_source": peak_src},
            "phases_ms": {"fwd_lookup+exchange": fwd_ms, "bwd_exchange": bwd_ms, "index+update": upd_ms,
                          "step_not_pipelined": serial_ms, "step_pipelined": ms_per_step},
            "nvlink": {"bytes_sent_per_rank_per_direction": a2a_bytes, "peak_gbs": 770.0,
                       "note": "phase times include the lookup / pack kernels; see profiles/ for the split"},
            "cpu_baseline": None,
        }
        emit_line(line)
    ens.close()
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
    ap.add_argument("--exchange", default=os.environ.get("ETB_EXCHANGE", "store"), choices=["store", "copy"],
                    help="N>1, fused exchange: 'store' = the lookup / scatter kernels store into peer HBM; 'copy' = they write "
                         "local staging blocks that the copy engines push over NVLink beside the next lookup")
    ap.add_argument("--index-beside", default=os.environ.get("ETB_INDEX_BESIDE", "forward"), choices=["forward", "exchange"],
                    help="N>1: run index! (side stream) beside the forward lookup or beside the backward exchange")
    ap.add_argument("--nccl-barrier", action="store_true",
                    help="N>1, fused exchange: a one-element NCCL all-reduce as barrier instead of the peer-memory flags")
    ap.add_argument("--no-self-check", action="store_true", help="N>1: skip the sharded-vs-single-GPU check before timing")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
