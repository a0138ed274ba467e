#!/bin/bash
# round 2, call 55 (2 GPUs): multi-device tests and torchrun parity on the final library
mkdir -p gpurun_out
export SB200_EXCHANGE_TIMEOUT_S=60
timeout -k 10 400 python -m pytest tests/test_sharded_capi_gpu.py tests/test_exchange_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu55.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu55.log
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/shard_check.py > gpurun_out/shard_check_n2e.log 2>&1
echo "shard_check rc=$?"; tail -1 gpurun_out/shard_check_n2e.log | cut -c1-200
