#!/bin/bash
# round 2, call 56: block cache on / off on ONE box: twelve one-shot C2 transposes each (final library)
mkdir -p gpurun_out
SB200_CACHE_MB=0 timeout -k 5 40 python tools/e2e_transpose_probe.py --reps 12 > gpurun_out/oneshot_cache_off.log 2>&1
echo "cache off:"; grep "^rep" gpurun_out/oneshot_cache_off.log | awk '{printf "%s ", $4} END {print ""}'
timeout -k 5 40 python tools/e2e_transpose_probe.py --reps 12 > gpurun_out/oneshot_cache_on.log 2>&1
echo "cache on:"; grep "^rep" gpurun_out/oneshot_cache_on.log | awk '{printf "%s ", $4} END {print ""}'
