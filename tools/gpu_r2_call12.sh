#!/bin/bash
# round 2, call 12 (2 GPUs): bench at N=2 and N=1 with the pipelined exchange tail; exchange micro-benchmark
mkdir -p gpurun_out
export SB200_EXCHANGE_TIMEOUT_S=60
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_d.json 2> gpurun_out/bench_n2_d.err
echo "bench n2 rc=$?"; tail -2 gpurun_out/bench_n2_d.err
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_d.json 2> gpurun_out/bench_n1_d.err
echo "bench n1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1_d.json", "gpurun_out/bench_n2_d.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"])
        for k, v in d["roofline_by_op"].items():
            print("   ", k, round(v["ms_per_launch"], 4), round(v["frac"], 3))
        print("    c4_strong", (d.get("c4_strong") or {}).get("speedup_vs_n1"))
    except Exception as e:
        print(f, "unreadable", e)
PY
