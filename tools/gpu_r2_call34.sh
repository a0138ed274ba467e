#!/bin/bash
# round 2, call 34: two-split transpose with as many scan threads as keys allow: parity + timings
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu34.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu34.log
OUT=gpurun_out/opbench34.jsonl; : > $OUT; : > gpurun_out/opbench34.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 5 --tag $tag >> $OUT 2>> gpurun_out/opbench34.err; }
trun scanw C2 transpose
trun scanw C4 transpose
trun scanw C3 transpose SB200_TRANSPOSE_PATH=split
trun scanw_p2_512 C3 transpose SB200_TRANSPOSE_PATH=split SB200_SPLIT_CFG2=512x8
trun scanw uniform:100000:100000:0.01:5 transpose
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep "trace" gpurun_out/opbench34.err | grep cached | sed 's/.*splits) //' | awk 'NR%5==0'
