#!/bin/bash
# round 2, call 11: transpose placement — column splits (L2 locality of neighbouring bands) and band counts
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu11.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu11.log
OUT=gpurun_out/opbench11.jsonl; : > $OUT
trun() { local tag=$1; shift; local wl=$1; shift
  env "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops transpose --reps 6 --tag $tag >> $OUT 2>> gpurun_out/opbench11.err; }
trun s4 C3
trun s16 C3 SB200_TRANSPOSE_SPLITS=16
trun s64 C3 SB200_TRANSPOSE_SPLITS=64
trun s256 C3 SB200_TRANSPOSE_SPLITS=256
trun b148s64 C3 SB200_TRANSPOSE_BANDS=148 SB200_TRANSPOSE_SPLITS=64
trun b148s256 C3 SB200_TRANSPOSE_BANDS=148 SB200_TRANSPOSE_SPLITS=256
trun b100s256 C3 SB200_TRANSPOSE_BANDS=100 SB200_TRANSPOSE_SPLITS=256
trun b200s128_k2 C3 SB200_TRANSPOSE_BANDS=200 SB200_TRANSPOSE_SPLITS=128 SB200_TRANSPOSE_KCOLS=2
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
