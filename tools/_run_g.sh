O=gpurun_out/g3; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --master-port 29811 tests/dist_gpu_check.py > $O/dist_check_n2.log 2>&1; tail -1 $O/dist_check_n2.log
timeout -s KILL 200 $TR --master-port 29812 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
timeout -s KILL 200 $TR --master-port 29813 bench.py --gpus 2 --steps 20 --warmup 5 --nccl-a2a --no-self-check > $O/bench_n2_nccl.json 2>> $O/bench_n2.err
python - <<P
import json
for f in ("bench_n2","bench_n2_nccl"):
    try:
        d=json.loads(open("gpurun_out/g3/%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],2), d.get("self_check"), d.get("phases_ms"), d["clocks"])
    except Exception as e: print(f,"ERR",e)
P
tail -2 $O/bench_n2.err
