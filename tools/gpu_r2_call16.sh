#!/bin/bash
# round 2, call 16: the two-split transpose of tall matrices: parity, then timings on C2 / C4 (and forced on C3, C1)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu16.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu16.log
OUT=gpurun_out/opbench16.jsonl; : > $OUT; : > gpurun_out/opbench16.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 6 --tag $tag >> $OUT 2>> gpurun_out/opbench16.err; }
trun split C2 transpose
trun banded C2 transpose SB200_TRANSPOSE_PATH=banded
trun split_sh9 C2 transpose SB200_SPLIT_SHIFT=9
trun split_sh11 C2 transpose SB200_SPLIT_SHIFT=11
trun split C4 transpose
trun split_sh9 C4 transpose SB200_SPLIT_SHIFT=9
trun split C3 transpose SB200_TRANSPOSE_PATH=split
trun split C1 transpose SB200_TRANSPOSE_PATH=split
trun default C1 transpose
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep "trace" gpurun_out/opbench16.err | sort | uniq -c | sort -rn | head -40
