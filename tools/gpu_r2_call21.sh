#!/bin/bash
# round 2, call 21: two-split transpose: CTA geometry sweep
mkdir -p gpurun_out
OUT=gpurun_out/opbench21.jsonl; : > $OUT; : > gpurun_out/opbench21.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 5 --tag $tag >> $OUT 2>> gpurun_out/opbench21.err; }
for cfg in 512x8 256x8 1024x8 256x16 1024x4; do
  trun $cfg C2 transpose SB200_SPLIT_CFG=$cfg
  trun ${cfg}_sh9 C2 transpose SB200_SPLIT_CFG=$cfg SB200_SPLIT_SHIFT=9
done
for cfg in 256x8 1024x4; do
  trun $cfg C3 transpose SB200_SPLIT_CFG=$cfg SB200_TRANSPOSE_PATH=split
done
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep "trace" gpurun_out/opbench21.err | grep cached | sed 's/.*splits) //' | awk 'NR%5==0'
