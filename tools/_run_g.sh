O=gpurun_out/g1; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 200 $TR --master-port 29812 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
ETB_E2E_DUPLEX=0 timeout -s KILL 200 $TR --master-port 29813 bench.py --gpus 2 --steps 10 --warmup 3 --no-self-check > $O/bench_n2_noduplex.json 2>> $O/bench_n2.err
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dist zipf > $O/zipf.json 2>$O/zipf.err
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/uni.json 2>$O/uni.err
timeout -s KILL 200 python tools/index_bench.py --out $O/index.jsonl > /dev/null 2>&1
python - <<P
import json
for f in ("bench_n2","bench_n2_noduplex","zipf","uni"):
    try:
        d=json.loads(open("gpurun_out/g1/%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],2), d.get("self_check"), d.get("phases_ms") or {k:round(v["ms"],3) for k,v in d["kernels"].items()})
    except Exception as e: print(f,"ERR",e)
for l in open("gpurun_out/g1/index.jsonl"):
    d=json.loads(l); print(d["case"], round(d["index_ms"],4), d["launches"], d["ok"])
P
tail -3 $O/bench_n2.err
