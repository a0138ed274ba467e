#!/bin/bash
# round 2, call 46 (2 GPUs): the multi-GPU paths on top of the block cache: GPU tests that need 2 devices, torchrun parity, bench --gpus 2
mkdir -p gpurun_out
export SB200_EXCHANGE_TIMEOUT_S=60
timeout -k 10 900 python -m pytest tests/test_sharded_capi_gpu.py tests/test_exchange_gpu.py tests/test_dropin_cpp.py -m gpu -x -q > gpurun_out/pytest_gpu46.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu46.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29631 tools/shard_check.py > gpurun_out/shard_check_n2d.log 2>&1
echo "shard_check rc=$?"; grep -c "sharded parity ok" gpurun_out/shard_check_n2d.log
timeout -k 10 900 $TR --master-port 29632 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_cache.json 2> gpurun_out/bench_n2_cache.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2_cache.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["n_gpus"], d.get("parity", {}).get("parity_checked"))
print(json.dumps(d.get("c4_strong"))[:600])
PY
