#!/bin/bash
set +e
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-1500; tail -5 gpurun_out/bench_n$N.err
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 2> /dev/null | tail -1 | cut -c1-300
echo "== sharded parity on $N ranks (NCCL)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/shard_check.py 2>&1 | tail -6
