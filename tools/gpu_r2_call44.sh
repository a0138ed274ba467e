#!/bin/bash
# round 2, call 44 (2 GPUs): phases of the one-process sharded step under trace
mkdir -p gpurun_out
SB200_TRACE=1 timeout -k 10 600 python tools/sharded_e2e.py --steps 2 > gpurun_out/sharded_e2e_trace.log 2>&1
echo "rc=$?"; grep -n "^gpus" gpurun_out/sharded_e2e_trace.log
grep -n "sharded create\|sharded destroy\|\] create:\|took" gpurun_out/sharded_e2e_trace.log | sed -n 1,200p | tail -90
