#!/bin/bash
# round 2, call 6: full GPU suite after the header/Matrix semantics change + band sweep v4 timings
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu6.log
OUT=gpurun_out/opbench6.jsonl; : > $OUT
run() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 10 --bmc 1 --tag $tag >> $OUT 2>> gpurun_out/opbench6.err; }
run v4 C4 spmv_t,spmv
run v4_cap8 C4 spmv_t,spmv SB200_BS_CAP=8
run v4_ahead C4 spmv_t,spmv SB200_BS_AHEAD=12288
run v4_640x3 C4 spmv_t,spmv SB200_BS_CFG=640,3
run v4 C2 spmv_t,spmv
run v4_ahead C2 spmv_t,spmv SB200_BS_AHEAD=12288
run v4_640x3 C2 spmv_t,spmv SB200_BS_CFG=640,3
run v4 C3 spmv_t,spmv
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/plain_ncu_target6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bandsweep -s 2 -c 1 -o gpurun_out/prof_bandsweep4_c2 \
  python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/ncu_bandsweep4.log 2>&1
echo "ncu rc=$?"
