#!/bin/bash
# multi-GPU evidence on N GPUs of one box: sharded parity over both exchanges, the bench with the library's
# peer-memory exchange and with NCCL, the exchange kernels alone, the reference arm under torchrun
set +e
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { python -c "
import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','row_path','gpu_launches')}); print({k:round(v['ms'],4) for k,v in d['per_op'].items()}); print(str(d.get('exchange'))[:90])"; }
echo "== sharded parity on $N ranks"
timeout 600 $TR --master-port 29513 tools/shard_check.py > gpurun_out/shard_check_n$N.log 2>&1; echo "rc=$?"; grep -c "sharded parity ok" gpurun_out/shard_check_n$N.log; grep "rank 0/" gpurun_out/shard_check_n$N.log | cut -c1-260
echo "== bench N=$N (peer-memory exchange kernels)"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; show gpurun_out/bench_n$N.json
echo "== bench N=$N --exchange nccl"
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 200 --warmup 20 --exchange nccl > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "rc=$?"; show gpurun_out/bench_n${N}_nccl.json
echo "== exchange kernels alone"
timeout 300 $TR --master-port 29517 tools/xg_bench.py 2>&1 | tail -1 | tee gpurun_out/xg_bench_n$N.json
if [ "${2:-}" = "ref" ]; then
echo "== reference arm under torchrun"
timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 2> /dev/null | tail -1 | cut -c1-300
fi
