#!/bin/bash
# round 2, call 5 (2 GPUs): sharded C ABI parity at n_gpus = 1, 2; two-rank exchange test; ncu of the band sweep at C2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus5.txt
timeout -k 10 900 python -m pytest tests/test_sharded_capi_gpu.py tests/test_exchange_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu5.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py > gpurun_out/shard_check_n2.log 2>&1
echo "shard_check rc=$?"; tail -3 gpurun_out/shard_check_n2.log
python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/plain_ncu_target5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bandsweep -s 2 -c 1 -o gpurun_out/prof_bandsweep3_c2 \
  python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/ncu_bandsweep3.log 2>&1
echo "ncu rc=$?"
