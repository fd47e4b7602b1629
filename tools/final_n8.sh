#!/bin/bash
# Round-2 measurement pass on 8 B200s of one box (gpurun --gpus 8): what the host <-> device path delivers with 8 ranks,
# the sharded path against the ORACLE, the bench line (default exchange and the copy-engine variant).
set -u
O=gpurun_out/final8; mkdir -p $O
T="timeout -s KILL"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T 200 $TR --master-port 29811 tools/pcie_probe.py 2>/dev/null > $O/r2_pcie_probe_n8.json
$T 300 $TR --master-port 29812 tests/dist_gpu_check.py > $O/r2_dist_check_n8.log 2>&1
$T 300 $TR --master-port 29813 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_bench_c2_n8.json 2> $O/r2_bench_c2_n8_selfcheck.log
$T 300 $TR --master-port 29814 bench.py --gpus 8 --steps 20 --warmup 5 --exchange copy --no-self-check > $O/r2_bench_c2_n8_copy_engines.json 2> $O/n8_copy.err
tail -2 $O/r2_dist_check_n8.log; cat $O/r2_pcie_probe_n8.json; cut -c1-600 $O/r2_bench_c2_n8.json; ls -la $O
