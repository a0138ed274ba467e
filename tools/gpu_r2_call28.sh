#!/bin/bash
# round 2, call 28 (8 GPUs): final bench at N=8 and N=2 (pipelined exchange tail, 74 exchange CTAs), sharded parity incl. the
# pushed sharded transpose on 8 ranks
mkdir -p gpurun_out
export SB200_EXCHANGE_TIMEOUT_S=120
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_f.json 2> gpurun_out/bench_n8_f.err
echo "bench n8 rc=$?"; tail -2 gpurun_out/bench_n8_f.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29583 tools/shard_check.py > gpurun_out/shard_check_n8_f.log 2>&1
echo "shard_check rc=$?"; tail -2 gpurun_out/shard_check_n8_f.log | cut -c1-300
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4_f.json 2> gpurun_out/bench_n4_f.err
echo "bench n4 rc=$?"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29584 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_f.json 2> gpurun_out/bench_n2_f.err
echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_n8_f.json", "gpurun_out/bench_n4_f.json", "gpurun_out/bench_n2_f.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"])
        for k, v in d["roofline_by_op"].items():
            print("   ", k, round(v["ms_per_launch"], 4), round(v["frac"], 3))
        print("    c4_strong", d.get("c4_strong", {}).get("speedup_vs_n1"))
    except Exception as e:
        print(f, "unreadable", e)
PY
