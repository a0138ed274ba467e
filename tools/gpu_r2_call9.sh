#!/bin/bash
# round 2, call 9 (8 GPUs): bench at N=8 (weak C2 + C4 strong scaling + in-run parity), sharded C ABI on 1/2/4/8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/gpus9.txt
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 rc=$?"; tail -3 gpurun_out/bench_n8.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench n4 rc=$?"
timeout -k 10 600 python -m pytest tests/test_sharded_capi_gpu.py tests/test_dropin_cpp.py -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu9.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tools/shard_check.py > gpurun_out/shard_check_n8.log 2>&1
echo "shard_check rc=$?"; tail -2 gpurun_out/shard_check_n8.log | cut -c1-300
python - <<'PY'
import json
for f in ("gpurun_out/bench_n8.json", "gpurun_out/bench_n4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"])
        for k, v in d["roofline_by_op"].items():
            print("   ", k, round(v["ms_per_launch"], 4), round(v["frac"], 3))
        print("    c4_strong", d.get("c4_strong"))
        print("    sections", {k: (v.get("error") if isinstance(v, dict) else v) for k, v in d.get("sections", {}).items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
