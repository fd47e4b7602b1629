#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): bench lines first (never under a profiler), then the ncu
# launch list and the --set full captures of the same commands.  Everything lands in gpurun_out/final/.
set -u
O=gpurun_out/final; mkdir -p $O
T="timeout -s KILL"
$T 900 python bench.py --steps 20 --warmup 5 2> $O/bench_c2_uniform.err > $O/r2_bench_c2_uniform.json
$T 600 python bench.py --steps 20 --warmup 5 --dist zipf --no-cpu-baseline 2>> $O/bench_c2_uniform.err > $O/r2_bench_c2_zipf_strict.json
ETB_E2E_DUPLEX=0 $T 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>> $O/bench_c2_uniform.err > $O/r2_bench_c2_uniform_e2e_serial.json
$T 600 python tools/bench_configs.py --config c1 --out $O/r2_c1_gather_update.jsonl > /dev/null 2>&1
$T 600 python tools/bench_configs.py --config c3 --out $O/r2_c3_zipf_update.jsonl > /dev/null 2>&1
ETB_STRICT_SLICED=0 $T 600 python tools/bench_configs.py --config c3 --out $O/r2_c3_zipf_update_unsliced.jsonl > /dev/null 2>&1
$T 600 python tools/bench_configs.py --config update --out $O/r2_update_sweep.jsonl > /dev/null 2>&1
$T 600 python tools/bench_configs.py --config c4local --out $O/r2_c4_local_split_tables.jsonl > /dev/null 2>&1
$T 900 python tools/bench_configs.py --config c5quick --out $O/r2_c5_quick.jsonl > /dev/null 2>&1
$T 300 python tools/index_bench.py --out $O/r2_index.jsonl > /dev/null 2>&1
ETB_IX_RANK=0 $T 300 python tools/index_bench.py --cases c2,c3 --out $O/r2_index_rank0.jsonl > /dev/null 2>&1
$T 600 python tools/bench_cached.py --out $O/r2_cached_table.jsonl > /dev/null 2>&1
$T 300 python tools/pcie_probe.py > $O/r2_pcie_probe_n1.json 2>/dev/null
$T 200 tools/ubench_rows > $O/r2_ubench_rows.jsonl 2>/dev/null
$T 100 tools/ubench_chain > $O/r2_ubench_chain.txt 2>/dev/null
# profiler passes (numbers printed under ncu are never bench values)
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_c3_strict.csv python tools/c3_once.py > /dev/null 2>&1
for k in pooled_kernel sgd_update_exact_kernel ix_scatter_kernel ix_write_records_kernel; do
  $T 600 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -s 2 -o $O/r2_full_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-overlap > /dev/null 2>&1
  ncu -i $O/r2_full_$k.ncu-rep --page raw --csv > $O/r2_full_$k.raw.csv 2>/dev/null
  rm -f $O/r2_full_$k.ncu-rep
done
k=long_strict_sliced_kernel
$T 300 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -s 1 -o $O/r2_full_$k python tools/c3_once.py > /dev/null 2>&1
ncu -i $O/r2_full_$k.ncu-rep --page raw --csv > $O/r2_full_$k.raw.csv 2>/dev/null
rm -f $O/r2_full_$k.ncu-rep
cuobjdump -sass embeddingtables.jl_b200/lib/libembtab_b200.so | grep -c "UBLKCP\|LDGSTS\|SYNCS" > $O/sass_counts.txt
ls -la $O
