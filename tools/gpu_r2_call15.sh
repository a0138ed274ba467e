#!/bin/bash
# round 2, call 15 (2 GPUs): sharded transpose over P2P pushes (shard_check on 2 ranks), exchange tests
mkdir -p gpurun_out
export SB200_EXCHANGE_TIMEOUT_S=60
timeout -k 10 600 python -m pytest tests/test_exchange_gpu.py tests/test_sharded_capi_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu15.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu15.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/shard_check.py > gpurun_out/shard_check_n2b.log 2>&1
echo "shard_check rc=$?"; tail -3 gpurun_out/shard_check_n2b.log | cut -c1-300
