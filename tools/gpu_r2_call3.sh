#!/bin/bash
# round 2, call 3: piece-based band sweep + chunk-sort transpose: parity subset, timings, variants, one ncu capture
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "band_companion or transpose" > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu3.log
OUT=gpurun_out/opbench3.jsonl; : > $OUT
run() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 10 --bmc 1 --tag $tag >> $OUT 2>> gpurun_out/opbench3.err; }
run cap16 C4 spmv_t,spmv
run cap8 C4 spmv_t,spmv SB200_BS_CAP=8
run cap16_ahead12k C4 spmv_t,spmv SB200_BS_AHEAD=12288
run cap16_640x3 C4 spmv_t,spmv SB200_BS_CFG=640,3
run cap16_512x4 C4 spmv_t,spmv SB200_BS_CFG=512,4
run cap16 C2 spmv_t,spmv
run cap8 C2 spmv_t,spmv SB200_BS_CAP=8
run cap16_ahead12k C2 spmv_t,spmv SB200_BS_AHEAD=12288
run cap8 C3 spmv_t,spmv
run cap16 C3 spmv_t,spmv SB200_BS_CAP=16
trun() { local tag=$1; shift; local wl=$1; shift
  env SB200_TRACE=1 "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops transpose --reps 8 --tag $tag >> $OUT 2>> gpurun_out/opbench3_trace.err; }
trun place_default C3
trun place_k1 C3 SB200_TRANSPOSE_KCOLS=1
trun place_k3 C3 SB200_TRANSPOSE_KCOLS=3
trun place_k4 C3 SB200_TRANSPOSE_KCOLS=4
trun place_512 C3 SB200_TRANSPOSE_CFG=512x3072
trun place_b444 C3 SB200_TRANSPOSE_BANDS=444
trun place_b200s8 C3 SB200_TRANSPOSE_BANDS=200 SB200_TRANSPOSE_SPLITS=8
trun banded C3 SB200_TRANSPOSE_PATH=banded
trun default C2
trun default C1
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep "trace" gpurun_out/opbench3_trace.err | sort | uniq -c | sort -rn | head -30
python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/plain_ncu_target.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bandsweep -s 2 -c 1 -o gpurun_out/prof_bandsweep2_c2 \
  python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/ncu_bandsweep2.log 2>&1
echo "ncu rc=$?"
