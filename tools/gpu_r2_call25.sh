#!/bin/bash
# round 2, call 25: two-split transpose: pass-2 geometry, ncu of the pipelined 16-byte-record version
mkdir -p gpurun_out
OUT=gpurun_out/opbench25.jsonl; : > $OUT; : > gpurun_out/opbench25.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 5 --tag $tag >> $OUT 2>> gpurun_out/opbench25.err; }
trun p2_1024 C2 transpose SB200_SPLIT_CFG2=1024x8
trun p2_1024_seg8 C2 transpose SB200_SPLIT_CFG2=1024x8 SB200_SPLIT_SEG=8
trun p2_1024 C4 transpose SB200_SPLIT_CFG2=1024x8
trun p2_1024_seg8 C4 transpose SB200_SPLIT_CFG2=1024x8 SB200_SPLIT_SEG=8
grep "trace" gpurun_out/opbench25.err | grep cached | sed 's/.*splits) //' | awk 'NR%5==0'
ncu --set full --clock-control none --import-source on -k regex:split_kernel -s 2 -c 2 -o gpurun_out/prof_split4_c2 \
  python tools/opbench.py --workload C2 --ops transpose --reps 3 > gpurun_out/ncu_split4.log 2>&1
echo "ncu rc=$?"
