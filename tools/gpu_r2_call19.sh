#!/bin/bash
# round 2, call 19: two-split transpose: segment length x band width sweep (L2 footprint of the pass-2 fronts)
mkdir -p gpurun_out
OUT=gpurun_out/opbench19.jsonl; : > $OUT; : > gpurun_out/opbench19.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 5 --tag $tag >> $OUT 2>> gpurun_out/opbench19.err; }
for seg in 1 4 32; do for sh in 9 10 11 12; do
  trun seg${seg}_sh${sh} C2 transpose SB200_SPLIT_SEG=$seg SB200_SPLIT_SHIFT=$sh
done; done
for seg in 1 4; do for sh in 10 11 12; do
  trun seg${seg}_sh${sh} C4 transpose SB200_SPLIT_SEG=$seg SB200_SPLIT_SHIFT=$sh
done; done
trun seg4 C3 transpose SB200_TRANSPOSE_PATH=split
trun seg4_sh9 C3 transpose SB200_TRANSPOSE_PATH=split SB200_SPLIT_SHIFT=9
trun seg4_sh7 C3 transpose SB200_TRANSPOSE_PATH=split SB200_SPLIT_SHIFT=7
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep "trace" gpurun_out/opbench19.err | grep cached | sed 's/.*splits) //' | awk 'NR%5==0'
