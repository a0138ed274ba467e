#!/bin/bash
# round 2, call 23: two-split transpose: do evict_last store hints keep the partially written sectors in L2?
mkdir -p gpurun_out
OUT=gpurun_out/opbench23.jsonl; : > $OUT; : > gpurun_out/opbench23.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 4 --tag $tag >> $OUT 2>> gpurun_out/opbench23.err; }
trun normal C2 transpose
trun hint C2 transpose SB200_SPLIT_DEBUG=2
trun hint_seg32 C2 transpose SB200_SPLIT_DEBUG=2 SB200_SPLIT_SEG=32
trun normal C4 transpose
trun hint C4 transpose SB200_SPLIT_DEBUG=2
grep "trace" gpurun_out/opbench23.err | grep cached | sed 's/.*splits) //' | awk 'NR%4==0'
