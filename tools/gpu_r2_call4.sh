#!/bin/bash
# round 2, call 4: band sweep with power-of-two pieces + long-run list; ncu of the transpose placement kernel
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "band_companion" > gpurun_out/pytest_gpu4.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu4.log
OUT=gpurun_out/opbench4.jsonl; : > $OUT
run() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 10 --bmc 1 --tag $tag >> $OUT 2>> gpurun_out/opbench4.err; }
run cap16 C4 spmv_t,spmv
run cap8 C4 spmv_t,spmv SB200_BS_CAP=8
run cap16_ahead12k C4 spmv_t,spmv SB200_BS_AHEAD=12288
run cap16_640x3 C4 spmv_t,spmv SB200_BS_CFG=640,3
run cap16 C2 spmv_t,spmv
run cap8 C2 spmv_t,spmv SB200_BS_CAP=8
run cap8 C3 spmv_t,spmv
run cap16 C3 spmv_t,spmv SB200_BS_CAP=16
run cap16_rows8k C4 spmv_t,spmv SB200_BMC_ROWS=8192
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep build $OUT | head -6
python tools/opbench.py --workload C3 --scale 0.1 --ops transpose --reps 3 > gpurun_out/plain_ncu_target4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:transpose_place -s 1 -c 1 -o gpurun_out/prof_place_c3s \
  python tools/opbench.py --workload C3 --scale 0.1 --ops transpose --reps 3 > gpurun_out/ncu_place.log 2>&1
echo "ncu rc=$?"
