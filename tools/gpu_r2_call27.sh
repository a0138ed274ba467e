#!/bin/bash
# round 2, call 27: first-call cost of a transpose (plan + scratch + kernels) in a warm process, both tall paths
mkdir -p gpurun_out
: > gpurun_out/first_call.jsonl
for wl in C2 C4; do
  timeout -k 10 600 python tools/first_call_probe.py --workload $wl --tag split >> gpurun_out/first_call.jsonl 2>> gpurun_out/first_call.err
  SB200_TRANSPOSE_PATH=banded timeout -k 10 600 python tools/first_call_probe.py --workload $wl --tag banded >> gpurun_out/first_call.jsonl 2>> gpurun_out/first_call.err
done
SB200_TRACE=1 timeout -k 10 600 python tools/first_call_probe.py --workload C2 --tag split_trace > /dev/null 2> gpurun_out/first_call_trace.err
cat gpurun_out/first_call.jsonl
grep trace gpurun_out/first_call_trace.err | sed 's/.*splits) //'
