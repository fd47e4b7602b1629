#!/usr/bin/env python
"""profiles/r2_ncu_kernels.json from the `ncu --page raw --csv` exports that tools/final_n1.sh leaves in gpurun_out/final/
(one capture per kernel): per-launch time, DRAM bytes, registers, occupancy, issue utilisation.  bench.py reads
`roofline.traffic` from this file."""
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
         "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
WANT = {"time_us": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "registers": "launch__registers_per_thread", "warps_active_per_sm": "sm__warps_active.avg.per_cycle_active",
        "warp_instructions": "smsp__inst_executed.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l2_hit_pct": "lts__t_sector_hit_rate.pct", "grid": "launch__grid_size",
        "occ_limit_smem_blocks": "launch__occupancy_limit_shared_mem", "occ_limit_regs_blocks": "launch__occupancy_limit_registers",
        "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"}


def one(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 20]
    head, units, vals = rows[0], rows[1], rows[-1]
    col = {n: i for i, n in enumerate(head)}
    out = {"kernel_name": vals[col["Kernel Name"]].replace("etb::", "")}
    for k, m in WANT.items():
        if m not in col:
            continue
        v = float(vals[col[m]].replace(",", ""))
        out[k] = v * SCALE.get(units[col[m]], 1.0)
    out["dram_bytes"] = out.get("dram_read", 0.0) + out.get("dram_write", 0.0)
    return out


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "final")
    kernels = {}
    for p in sorted(glob.glob(os.path.join(src, "r2_full_*.raw.csv"))):
        name = os.path.basename(p)[len("r2_full_"):-len(".raw.csv")]
        try:
            kernels[name] = one(p)
        except Exception as e:  # an empty export (the capture found no such kernel)
            print("skipped", p, e)
    doc = {"source": "ncu --set full --clock-control none -k regex:<kernel> -c 1 -s 2  python bench.py --steps 1 --warmup 3 "
                     "--no-cpu-baseline --no-overlap (long_strict_sliced_kernel: tools/c3_once.py); round 2, tools/final_n1.sh; per launch",
           "kernels": kernels}
    with open(os.path.join(ROOT, "profiles", "r2_ncu_kernels.json"), "w") as f:
        json.dump(doc, f, indent=1)
    for k, v in kernels.items():
        print(k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != "kernel_name"})


if __name__ == "__main__":
    main()
