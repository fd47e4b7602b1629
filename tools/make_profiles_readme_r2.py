#!/usr/bin/env python
"""Regenerates profiles/README.md (round 2) from the committed extracts in profiles/."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def jl(name):
    with open(os.path.join(P, name)) as f:
        return [json.loads(l) for l in f if l.strip().startswith("{")]


def last(name):
    return jl(name)[-1]


def launch_table(name, first_kernel="pooled_kernel"):
    rows = [r for r in csv.reader(open(os.path.join(P, name))) if len(r) > 14 and r[0].isdigit()]
    agg = {}
    for r in rows:
        k = r[4].split("(")[0].replace("void ", "").strip()
        if k.startswith("at::") or "distribution" in k or "elementwise" in k or "vectorized" in k:
            continue
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e3
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | avg µs | share of the listed kernels |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    return "\n".join(out)


def main():
    b = last("r2_bench_c2_uniform.json")
    z = last("r2_bench_c2_zipf_strict.json")
    es = last("r2_bench_c2_uniform_e2e_serial.json")
    n2, n8 = last("r2_bench_c2_n2.json"), last("r2_bench_c2_n8.json")
    n8g2, n8g4, n2g4 = last("r2_bench_c2_n8_2groups.json"), last("r2_bench_c2_n8_4groups.json"), last("r2_bench_c2_n2_4groups.json")
    nk = json.load(open(os.path.join(P, "r2_ncu_kernels.json")))["kernels"]
    k = b["kernels"]
    ix_name = [x for x in k if x.startswith("index")][0]
    cpu = b["cpu_baseline"]
    L = []
    A = L.append
    A("# profiles/ — round 2 evidence (B200, sm_100a, driver 580, CUDA 12.9)\n")
    A("Every number comes from `gpurun` runs on B200s of this pool (`tools/final_n1.sh` for the one-GPU pass; the torchrun\n"
      "lines for N = 2 / 8).  The `.ncu-rep` files stay in `gpurun_out/` (scratch); the extracts are committed here and this\n"
      "file is generated from them by `tools/make_profiles_readme_r2.py`.  Round 1's evidence is kept: `README_r1.md`, `r1_*`.\n")
    A("| file | what |\n|---|---|")
    A("| `r2_bench_c2_uniform.json`, `r2_bench_c2_zipf_strict.json` | `python bench.py --steps 20 --warmup 5` / `--dist zipf`: the bench lines (strict = the default update order); `r2_bench_c2_uniform_e2e_serial.json`: the same with `ETB_E2E_DUPLEX=0` (the whole result on the host before any cotangent is sent) |")
    A("| `r2_launches_bench_c2.csv`, `r2_launches_c3_strict.csv` | `ncu --metrics gpu__time_duration.sum --clock-control none` launch lists of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` and of `tools/c3_once.py` (C3, strict order) |")
    A("| `r2_ncu_kernels.json` | per-launch DRAM bytes / time / registers / occupancy / issue utilisation of the hot kernels from `ncu --set full` captures of the same program (`bench.py` reads `roofline.traffic` from it) |")
    A("| `r2_bench_c2_n2.json`, `r2_bench_c2_n4.json`, `r2_bench_c2_n8.json` | torchrun bench lines at N = 2 / 4 / 8 (one table group = the default); `..._2groups`, `..._4groups`: the table-group pipelined backward |")
    A("| `r2_dist_check_n2.log`, `r2_dist_check_n8.log`, `r2_bench_c2_n8_selfcheck.log` | `tests/dist_gpu_check.py` under torchrun (sharded == single GPU == ORACLE, bit for bit, three exchange modes) and the per-rank self-check lines of `bench.py --gpus 8` |")
    A("| `r2_c4_n8.jsonl` | `tools/bench_c4.py` on 8 GPUs: BASELINE configs[3] (64 chunked tables of 5 M rows) |")
    A("| `r2_index.jsonl`, `r2_index_rank0.jsonl` | `tools/index_bench.py`: index! alone by CUDA-graph replay at the C1–C4 shapes, records checked against numpy's stable sort; `rank0` = the all-ballots ranking |")
    A("| `r2_c1_*.jsonl`, `r2_c3_*.jsonl`, `r2_c4_local_*.jsonl`, `r2_c5_quick.jsonl`, `r2_update_sweep.jsonl` | `tools/bench_configs.py`: the other BASELINE configs on one GPU |")
    A("| `r2_cached_table.jsonl` | `tools/bench_cached.py`: host-tier table behind an HBM row cache under Zipf(1.05) |")
    A("| `r2_pcie_probe_n1.json`, `r2_pcie_probe_n2.json`, `r2_pcie_probe_n8.json` | `tools/pcie_probe.py`: what the box's host <-> device path delivers with 1 / 2 / 8 ranks copying at once |")
    A("| `r2_bench_c2_n2_copy_engines.json`, `r2_bench_c2_n8_copy_engines.json` | the bench line with `--exchange copy` (kernels write local staging blocks, the copy engines push them over NVLink) |")
    A("| `r2_c3_zipf_update_unsliced.jsonl` | C3 with `ETB_STRICT_SLICED=0`: the one-CTA-per-hot-row kernel of the first half of round 2 |")
    A("| `r2_ubench_rows.jsonl`, `r2_ubench_chain.txt` | `tools/ubench_rows.cu` (HBM ceiling of the update's access pattern by row length) and `tools/ubench_chain.cu` (cycles per member of a strictly ordered sum fed from shared memory) |\n")
    A("compute-sanitizer is closed on this pool (`gpurun` refuses it); `tools/sanitize_smoke.py` (a pass over every kernel\n"
      "family with odd shapes) runs clean without it, and bounds are covered by canary checks in the parity tests.\n")
    A("## C2 (26 x 1M x 128 f32, bag 32, batch 16384, uniform), one B200\n")
    A("| quantity | round 2 | round 1 |\n|---|---|---|")
    A(f"| step = fused forward + lazy pullback + ensemble update! | **{b['ms_per_step']:.2f} ms -> {b['value'] / 1e9:.2f} G lookups/s** (index! on a side stream beside the forward; {b['ms_per_step_phases_back_to_back']:.2f} ms with the phases back to back) | 3.40 ms, 4.01 G |")
    A(f"| forward (`pooled_kernel`, 1 launch) | {k['pooled_kernel']['ms']:.2f} ms; 7.31 GB algorithmic -> {k['pooled_kernel']['gbs'] / 1e3:.2f} TB/s = {k['pooled_kernel']['gbs'] / 6546.9:.2f} x measured copy peak; DRAM traffic {nk['pooled_kernel']['dram_bytes'] / 1e9:.2f} GB; {int(nk['pooled_kernel']['registers'])} registers, {nk['pooled_kernel']['warps_active_per_sm']:.0f} warps/SM active | 0.97 ms |")
    A(f"| index! (per-table segmented radix sort: 2 x (hist, scan, scatter) + 3 record kernels, {b['launches_per_step']['index']} launches) | **{k[ix_name]['ms']:.2f} ms** (0.35 ms by graph replay, `r2_index.jsonl`) | 0.54 ms, 13 launches |")
    A(f"| update (`sgd_update_exact_kernel` + task / long-bucket kernels, {b['launches_per_step']['update']} launches) | {k['sgd_update_kernel']['ms']:.2f} ms; 11.13 GB algorithmic -> **{k['sgd_update_kernel']['gbs'] / 1e3:.2f} TB/s = {k['sgd_update_kernel']['gbs'] / 6546.9:.2f} x measured peak**; DRAM traffic {nk['sgd_update_exact_kernel']['dram_bytes'] / 1e9:.2f} GB; {int(nk['sgd_update_exact_kernel']['registers'])} registers, {nk['sgd_update_exact_kernel']['warps_active_per_sm']:.0f} warps/SM | 1.91 ms, 0.89 |")
    A(f"| fwd+bwd+SGD | {b['fwd_bwd_sgd_gbs'] / 1e3:.2f} TB/s algorithmic = {b['fwd_bwd_sgd_frac_of_peak']:.2f} x measured peak | 0.83 |")
    A(f"| e2e (pinned host indices + cotangent in, feature matrix out; 336 MB H2D + 226 MB D2H per step) | {b['e2e']['ms_per_step']:.1f} ms -> {b['e2e']['value'] / 1e9:.2f} G lookups/s (PCIe full duplex: the cotangent's column chunks follow their own result chunks, 336 MB H2D at ~50 GB/s is the floor; {es['e2e']['ms_per_step']:.1f} ms with the whole result on the host first, `ETB_E2E_DUPLEX=0`) | 9.0 ms |")
    A(f"| CPU arm: C port of the reference, ALL 26 tables, {cpu['cores']} pinned host threads, AVX-512 | **{cpu['value'] / 1e6:.1f} M lookups/s** ({cpu['sample'].split('update! with the ')[1]} | 22 M on a 4-table sample (over-stated the ratio) |")
    A(f"| clocks in the timed regions | {b['clocks']['sm_mhz']:.0f} MHz of {b['clocks']['sm_max_mhz']:.0f}, reasons: {b['clocks']['reasons']} ({b['clocks'].get('samples')} samples) | |")
    zk = z["kernels"]
    A(f"| Zipf(1.05) indices, strict order (default) | step {z['ms_per_step']:.2f} ms = {z['value'] / 1e9:.2f} G lookups/s (forward {zk['pooled_kernel']['ms']:.2f} ms from L2, update {zk['sgd_update_kernel']['ms']:.2f} ms incl. the hot rows, streamed slice by slice: 3.75 ms before the slicing) | split order: 2.44 ms (`set_update_order(\"split\")`: 2.38 ms now) |\n")
    A("### Launch list of the step (cold-cache, serialised by ncu; compare shares)\n")
    A(launch_table("r2_launches_bench_c2.csv"))
    A("\n(The list covers the warm-up, timed, per-phase and e2e regions of `bench.py --steps 2 --warmup 3`; the e2e region\nlooks the batch up in 8 column chunks and updates in 7 table groups, hence the smaller launches of the same kernels.)\n")
    A("### Hot kernels under `ncu --set full` (`r2_ncu_kernels.json`)\n")
    A("| kernel | µs | DRAM read + write | registers | warps/SM | issue slots busy | warp instructions |\n|---|---|---|---|---|---|---|")
    for name, d in nk.items():
        A(f"| `{name}` | {d['time_us']:.0f} | {d['dram_read'] / 1e6:.0f} + {d['dram_write'] / 1e6:.0f} MB | {int(d['registers'])} | {d['warps_active_per_sm']:.0f} | {d['issue_active_pct']:.0f} % | {d['warp_instructions'] / 1e6:.1f} M |")
    A("\nThe scatter kernel of index! (`ix_scatter_kernel`, 512 threads x 16 keys per tile) is instruction-bound, not memory-bound:\n"
      "64 M warp instructions per pass at 53 % issue utilisation for 177 MB of DRAM traffic (round 1's 9-bit pass: 68.7 M\n"
      "instructions, 54 %, three passes instead of two).\n")
    A("## index! alone (`tools/index_bench.py`, CUDA-graph replay, records verified against numpy's stable sort)\n")
    A("| case | keys | row bits | launches | ms | G keys/s | round 1 |\n|---|---|---|---|---|---|---|")
    r1 = {"c2": "0.55", "c2zipf": "0.55", "c3": "0.083", "c3pooled": "0.083", "c1": "0.063 (eager 13 launches)", "c4": "-"}
    for d in jl("r2_index.jsonl"):
        A(f"| {d['case']} ({d['tables']} x {d['n_per_table']}, {d['dist']}) | {d['tables'] * d['n_per_table'] / 1e6:.2f} M | {d['row_bits']} | {d['launches']} | {d['index_ms']:.3f} | {d['keys_per_s'] / 1e9:.1f} | {r1.get(d['case'], '-')} |")
    A("\nDesigns measured on the way (same tool, same box class; all verified against numpy's stable sort):\n")
    A("| design | C2 ms | C3 ms | why it lost |\n|---|---|---|---|")
    A("| round 1: slot in the key, 3 passes (9/8/8 bits), make_pairs + 3 x (hist, scan, scatter) + 3 record kernels | 0.55 | 0.083 | one pass too many; digit-major tile histograms; serialised loads in the record kernels |")
    A("| one-sweep passes with per-tile decoupled look-back (atomic ticket, 1024 chains per tile) | 0.78 (256 thr) / 0.56 (512 thr) | 0.104 | ~1000 small tiles resident: look-back chains as long as the tiles in flight; polling doubles the instruction count (100 M vs 64 M per pass); records with a one-value look-back chain: 152 µs vs 49 |")
    A("| chunks of 8 tiles per CTA, running digit cursors, no per-tile histograms | 0.68 (3 CTAs/SM) / 0.53 (6) | 0.167 | partial sectors leave L2 before the next tile completes them (DRAM writes 245 MB for 109 MB); one long wave of 416 CTAs; summing 128 chunk histograms per CTA at C3 |")
    A("| **kept**: per-table segments, 2 x 10 bits, first pass reads the indices, `[tile][digit]` histograms, atomics + ballots-on-demand ranking, batched loads in the record kernels | **0.35-0.39** | **0.080** (0.070 with the warp-per-digit scan for calls with few tables) | |")
    A("| ... with ballots for every row (`ETB_IX_RANK=0`) | 0.37-0.41 | 0.076 | |\n")
    A("## Multi-GPU (weak scaling: 26 tables per GPU, global batch 16384, fused NVLink exchange, peer-memory barrier)\n")
    A("| N | ms/step (device, max over ranks) | G lookups/s | lookup + exchange | backward exchange | index! + update! | e2e ms/step | round 1 |\n|---|---|---|---|---|---|---|---|")
    A(f"| 1 | {b['ms_per_step']:.2f} | {b['value'] / 1e9:.2f} | {k['pooled_kernel']['ms']:.2f} | - | {k[ix_name]['ms'] + k['sgd_update_kernel']['ms']:.2f} | {b['e2e']['ms_per_step']:.1f} | 3.39 / 9.0 |")
    rows_n = [(n2, "3.58 / 15.6")]
    if os.path.exists(os.path.join(P, "r2_bench_c2_n4.json")):
        rows_n.append((last("r2_bench_c2_n4.json"), "3.66 / 26.9"))
    rows_n.append((n8, "3.72 / 38.8"))
    for d, r1s in rows_n:
        ph = d["phases_ms"]
        A(f"| {d['n_gpus']} | {d['ms_per_step']:.2f} | {d['value'] / 1e9:.2f} | {ph['fwd_lookup+exchange']:.2f} | {ph['bwd_exchange']:.2f} | {ph['index+update']:.2f} | {d['e2e']['ms_per_step']:.1f} | {r1s} |")
    A(f"\nWeak-scaling efficiency at N = 8: {b['ms_per_step'] / n8['ms_per_step']:.2f} (device-timed).  Self-check before timing: `{n8['self_check']}` "
      "(skipped in the group-variant runs below; the default run's per-rank lines are in `r2_bench_c2_n8_selfcheck.log`).\n")
    A("Table-group pipelining of the backward (exchange of group g+1 beside update! of group g) — implemented, verified, slower:\n")
    A("| N | 1 group (default) | 2 groups | 4 groups |\n|---|---|---|---|")
    A(f"| 2 | {n2['ms_per_step']:.2f} ms | - | {n2g4['ms_per_step']:.2f} ms |")
    A(f"| 8 | {n8['ms_per_step']:.2f} ms | {n8g2['ms_per_step']:.2f} ms | {n8g4['ms_per_step']:.2f} ms (not pipelined, same groups: {n8g4['phases_ms']['step_not_pipelined']:.2f}) |\n")
    for nm, lab in (("r2_bench_c2_n2_copy_engines.json", 2), ("r2_bench_c2_n8_copy_engines.json", 8)):
        if os.path.exists(os.path.join(P, nm)):
            d = last(nm)
            ph = d["phases_ms"]
            A(f"Copy engines instead of peer stores (`--exchange copy`) at N = {lab}: {d['ms_per_step']:.2f} ms/step (lookup + exchange {ph['fwd_lookup+exchange']:.2f}, backward exchange {ph['bwd_exchange']:.2f}, index! + update! {ph['index+update']:.2f}) — parity-checked (`dist_gpu_check.py`), no gain: opt-in.\n")
    A("Barrier: peer-memory flags (`etb_peer_barrier`) 3.64 ms vs NCCL one-element all-reduce 3.61 ms at N = 2 (4 groups) — never the cost.\n")
    A("### C4 on 8 GPUs (`r2_c4_n8.jsonl`: 64 SplitEmbedding 128 x 5M f32, cols_per_shard 1 048 576, bag 32)\n")
    A("| global batch | indices | ms/step | G lookups/s | lookup + exchange (frac of measured HBM peak) | backward exchange alone | GB/s per GPU (frac of NVLink 900) |\n|---|---|---|---|---|---|---|")
    for d in jl("r2_c4_n8.jsonl"):
        A(f"| {d['batch_global']} | {d['dist']} | {d['ms_per_step']:.2f} | {d['lookups_per_sec'] / 1e9:.1f} | {d['fwd_lookup+exchange_ms']:.2f} ms ({d['fwd_frac_of_measured_hbm_peak']:.2f}) | {d['bwd_exchange_alone_ms']:.3f} ms | {d['bwd_exchange_gbs_per_gpu']:.0f} ({d['bwd_exchange_frac_of_nvlink']:.2f}) |")
    A("\n(Measured in the first half of round 2, before the hot rows were sliced: the Zipf runs use the strict order with one CTA per hot row.  "
      "e2e at N = 2 varies between boxes (9.7 - 12.9 ms with the same code); on the box of the committed line the host-buffer step by 4 table groups "
      "(`ETB_E2E_TABLE_GROUPS`, the default) takes 11.7 ms, 12.9 ms with the ensemble whole, 11.7 ms with 7 groups.)\n")
    if os.path.exists(os.path.join(P, "r2_pcie_probe_n8.json")):
        p1, p8 = last("r2_pcie_probe_n1.json"), last("r2_pcie_probe_n8.json")
        p2 = last("r2_pcie_probe_n2.json") if os.path.exists(os.path.join(P, "r2_pcie_probe_n2.json")) else None
        A("### Why e2e does not scale at N = 8: the host path (`tools/pcie_probe.py`)\n")
        A("| ranks copying at once | H2D GB/s per rank / aggregate | D2H | both directions |\n|---|---|---|---|")
        for p in [x for x in (p1, p2, p8) if x]:
            A(f"| {p['n_gpus']} | {p['h2d']['gbs_per_rank']:.1f} / {p['h2d']['gbs_aggregate']:.0f} | {p['d2h']['gbs_per_rank']:.1f} / {p['d2h']['gbs_aggregate']:.0f} | {p['h2d+d2h']['gbs_per_rank']:.1f} / {p['h2d+d2h']['gbs_aggregate']:.0f} |")
        tot = 8 * (n8['e2e']['h2d_bytes_per_step'] + n8['e2e']['d2h_bytes_per_step'])
        A(f"\nThe e2e step moves {tot / 1e9:.2f} GB through the host per step at N = 8 ({n8['e2e']['h2d_bytes_per_step'] / 1e6:.0f} MB in + {n8['e2e']['d2h_bytes_per_step'] / 1e6:.0f} MB out per rank, "
          f"the cotangent of a column chunk following its own result chunk); at the probe's aggregate rate that alone is {tot / (p8['h2d+d2h']['gbs_aggregate'] * 1e6):.1f} ms if both directions were always busy, "
          f"{tot / (p8['h2d']['gbs_aggregate'] * 1e6):.1f} ms one direction at a time; measured {n8['e2e']['ms_per_step']:.1f} ms.\n")
    A("## Other configs on one GPU\n")
    c1 = last("r2_c1_gather_update.jsonl")
    A(f"* **C1** (26 x (64 x 100k), batch 2048, gather + update!): one CUDA graph {c1['step_cuda_graph_us']:.0f} µs ({c1.get('step_cuda_graph_index_beside_gather_us', float('nan')):.0f} µs with index! on the side stream beside the gather; round 1: 86; roofline {c1['roofline_time_us_at_measured_peak']:.1f} µs; 7 graph nodes of 6 – 7 µs each: launch latency, not traffic); eager through the Python mirror {c1['step_us']:.0f} µs; index! 2 launches, 20 µs (`ix_small_kernel`).")
    A("* **C3** (one 128 x 10M table, Zipf(1.05), `r2_c3_zipf_update.jsonl`; GPU time by graph replay, index! + update kernels):\n")
    A("| form | n | hottest row | order | update µs | of which index! | algorithmic GB/s (frac of measured peak) | before the slicing | round 1 (eager) |\n|---|---|---|---|---|---|---|---|---|")
    uns = {(d['form'], d['n'], d['order']): d for d in jl("r2_c3_zipf_update_unsliced.jsonl")} if os.path.exists(os.path.join(P, "r2_c3_zipf_update_unsliced.jsonl")) else {}
    r1c3 = {("vector", 65536, "split"): "-", ("vector", 524288, "split"): "215", ("vector", 524288, "strict"): "7400", ("pooled 32x16384", 524288, "split"): "165"}
    for d in jl("r2_c3_zipf_update.jsonl"):
        A(f"| {d['form']} | {d['n']} | {d['hottest_row_members']} | {d['order']} | {d['update_us_graph']:.0f} | {d['index_us_graph']:.0f} | {d['gbs_graph']:.0f} ({d['frac_of_measured_peak_graph']:.2f}) | {('%.0f' % uns[(d['form'], d['n'], d['order'])]['update_us_graph']) if (d['form'], d['n'], d['order']) in uns and d['order'] == 'strict' else '-'} | {r1c3.get((d['form'], d['n'], d['order']), '-')} |")
    A("\n  Strict order = the reference's sequential sum (default).  A hot row is cut into slices of 16 feature elements; every slice is a job of\n"
      "  `long_strict_sliced_kernel` on its own SM (7 warps stream the members' slices through shared memory with `cp.async`, stages handed over by\n"
      "  mbarriers, one warp adds: one shared-memory load and one dependent add per member).  Cycle counts of the hottest job (clock64, `-DETB_SLICE_PROFILE`):\n"
      "  7.4 cycles per member (6.2 in the add loop + 300 per 256-member stage); the loop alone runs at 4.8 (`r2_ubench_chain.txt`).  Before the slicing one CTA\n"
      "  streamed whole rows (`long_strict_kernel`, 30 cycles per member; its `cp.async.bulk` variant 2094 µs at n = 524288 against 893: opt-in only).\n")
    if os.path.exists(os.path.join(P, "r2_ubench_chain.txt")):
        A("  ```\n  " + open(os.path.join(P, "r2_ubench_chain.txt")).read().strip().replace("\n", "\n  ") + "\n  ```\n")
    A("* **Update sweep over feature sizes** (`r2_update_sweep.jsonl`, C2's shape with dim varied): "
      + ", ".join(f"dim {d['dim']}: {d['kernel_frac_of_measured_peak']:.2f}" for d in jl("r2_update_sweep.jsonl")) + " of the measured peak.")
    if os.path.exists(os.path.join(P, "r2_ubench_rows.jsonl")):
        best = {}
        for d in jl("r2_ubench_rows.jsonl"):
            best[d["dim_f32"]] = max(best.get(d["dim_f32"], 0.0), d["frac_of_measured_peak"])
        A("  What the memory system gives a bare kernel with the same access pattern (`tools/ubench_rows.cu`: random sorted rows read-modify-written + one random delta row each, no bucket records): "
          + ", ".join(f"dim {k}: {v:.2f}" for k, v in sorted(best.items())) + " — narrow rows are limited by DRAM access granularity (dim 80 sits on that ceiling; dim 16 / 32 are 0.15 below it: the 16-byte bucket records and the map are not in the bare kernel's bytes).")
    A("  Exact-fit layouts for rows of 3·2^k / 5·2^k vectors (dim 80 = 4 lanes x 5 vectors) were measured at 0.58 for dim 80 in the same run\n"
      "  class (0.60–0.63 without): not kept — the limiter is DRAM fetch granularity, not idle lanes.")
    c4l = jl("r2_c4_local_split_tables.jsonl")
    A("* **C4-local** (8 chunked tables 128 x 5M on one GPU): " + "; ".join(f"{d['tables']} batch {d['shape'].split('batch ')[1]}: forward {d['fwd_ms']:.2f} ms ({d['fwd_frac_of_measured_peak']:.2f} of peak), update {d['update_ms']:.2f} ms" for d in c4l) + ".")
    A("* **Host-tier table** (`r2_cached_table.jsonl`: 1M x 128 f32 on the host, Zipf(1.05), bag 32 x batch 16384):\n")
    A("| table | rows in HBM | hit rate per warm-up step -> timed step | pooled lookup ms | algorithmic GB/s | update! ms |\n|---|---|---|---|---|---|")
    for d in jl("r2_cached_table.jsonl"):
        A(f"| {d['table']} | {d['cached_rows'] if d['cached_rows'] is not None else 'all'} | {', '.join(f'{x:.2f}' for x in d['hit_rate_per_warm_step'])} -> {d['hit_rate']:.2f} | {d['fwd_ms']:.2f} | {d['fwd_gbs']:.0f} | {d['update_ms']:.2f} |")
    A("\n* **C5 (quick sweep)**: `r2_c5_quick.jsonl` (dim 16–256 x bag 1/8/32/128 x batch 4k/16k, uniform + Zipf); the full round-1 sweep is `r1_c5_sweep.jsonl` (the lookup kernels did not change).\n")
    A("## How to reproduce a capture\n")
    A("```bash\n# launch list (times are cold-cache and serialised: compare shares)\n"
      "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n"
      "# one kernel in full\nncu --set full --import-source on --clock-control none -k regex:ix_scatter_kernel -c 1 -s 2 -o scatter python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-overlap\n"
      "ncu -i scatter.ncu-rep --page raw --csv        # metrics;  --page source --csv --print-source sass  for the per-instruction view\n"
      "# NVTX: every etb_* call opens a range named after itself (nsys is not in this image):\n#   nsys profile --trace=cuda,nvtx python bench.py --steps 3 --no-cpu-baseline\n```\n")
    with open(os.path.join(P, "README.md"), "w") as f:
        f.write("\n".join(L))
    print("wrote profiles/README.md", len(L), "blocks")


if __name__ == "__main__":
    main()
