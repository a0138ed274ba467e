#!/bin/bash
# round 2, call 39: where the one-shot host-buffer transpose (e2e.transpose, 485 ms at C2) spends its time
mkdir -p gpurun_out
SB200_TRACE=1 timeout -k 10 600 python tools/e2e_transpose_probe.py --reps 4 > gpurun_out/e2e_transpose_probe.log 2>&1
echo "probe rc=$?"; grep -v "^\[sb200 trace\] create" gpurun_out/e2e_transpose_probe.log | tail -60
