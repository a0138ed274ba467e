#!/bin/bash
# round 2, call 43 (2 GPUs): one process over 1 and 2 GPUs end to end from pageable host arrays (sb200_sharded_*)
mkdir -p gpurun_out
timeout -k 10 600 python tools/sharded_e2e.py --steps 5 --out gpurun_out/sharded_e2e_n2.json > gpurun_out/sharded_e2e_n2.log 2>&1
echo "rc=$?"; grep "^gpus" gpurun_out/sharded_e2e_n2.log; tail -3 gpurun_out/sharded_e2e_n2.log | cut -c1-300
