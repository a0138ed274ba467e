#!/bin/bash
# round 2, call 24: two-split transpose with 16-byte records, one-chunk-ahead pipeline: parity + timings
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu24.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu24.log
OUT=gpurun_out/opbench24.jsonl; : > $OUT; : > gpurun_out/opbench24.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 5 --tag $tag >> $OUT 2>> gpurun_out/opbench24.err; }
trun rec16 C2 transpose
trun rec16_sh9 C2 transpose SB200_SPLIT_SHIFT=9
trun rec16_sh11 C2 transpose SB200_SPLIT_SHIFT=11
trun rec16_p2_512 C2 transpose SB200_SPLIT_CFG2=512x8
trun rec16 C4 transpose
trun rec16_sh9 C4 transpose SB200_SPLIT_SHIFT=9
trun rec16 C3 transpose SB200_TRANSPOSE_PATH=split
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep "trace" gpurun_out/opbench24.err | grep cached | sed 's/.*splits) //' | awk 'NR%5==0'
