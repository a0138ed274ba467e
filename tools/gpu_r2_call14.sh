#!/bin/bash
# round 2, call 14: L2 fetch granularity hint vs the sector-random reads of the transpose (and the banded kernels)
mkdir -p gpurun_out
OUT=gpurun_out/opbench14.jsonl; : > $OUT
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 6 --tag $tag >> $OUT 2>> gpurun_out/opbench14.err; }
trun default C3 transpose
trun l2f32 C3 transpose SB200_L2_FETCH=32
trun l2f128 C3 transpose SB200_L2_FETCH=128
trun default C2 transpose,rowSums,spmv_t
trun l2f32 C2 transpose,rowSums,spmv_t SB200_L2_FETCH=32
trun default C4 transpose
trun l2f32 C4 transpose SB200_L2_FETCH=32
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
