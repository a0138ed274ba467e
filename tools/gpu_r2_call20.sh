#!/bin/bash
# round 2, call 20: two-split transpose with per-pass fast ranks: parity, timings, ncu
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu20.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu20.log
OUT=gpurun_out/opbench20.jsonl; : > $OUT; : > gpurun_out/opbench20.err
trun() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 5 --tag $tag >> $OUT 2>> gpurun_out/opbench20.err; }
trun split C2 transpose
trun split_sh9 C2 transpose SB200_SPLIT_SHIFT=9
trun split_sh11 C2 transpose SB200_SPLIT_SHIFT=11
trun split C4 transpose
trun split C3 transpose SB200_TRANSPOSE_PATH=split
trun split C1 transpose SB200_TRANSPOSE_PATH=split
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep "trace" gpurun_out/opbench20.err | grep cached | sed 's/.*splits) //' | awk 'NR%5==0'
ncu --set full --clock-control none --import-source on -k regex:split_kernel -s 2 -c 2 -o gpurun_out/prof_split3_c2 \
  python tools/opbench.py --workload C2 --ops transpose --reps 3 > gpurun_out/ncu_split3.log 2>&1
echo "ncu rc=$?"
