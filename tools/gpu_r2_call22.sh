#!/bin/bash
# round 2, call 22: where does the two-split transpose lose time: scatter (memory side) or ranking (SM side)?
mkdir -p gpurun_out
OUT=gpurun_out/opbench22.jsonl; : > $OUT; : > gpurun_out/opbench22.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 4 --tag $tag >> $OUT 2>> gpurun_out/opbench22.err; }
trun normal C2 transpose
trun linear C2 transpose SB200_SPLIT_DEBUG=1
trun linear_sh9 C2 transpose SB200_SPLIT_DEBUG=1 SB200_SPLIT_SHIFT=9
trun linear_sh11 C2 transpose SB200_SPLIT_DEBUG=1 SB200_SPLIT_SHIFT=11
grep "trace" gpurun_out/opbench22.err | grep cached | sed 's/.*splits) //' | awk 'NR%4==0'
