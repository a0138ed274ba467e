#!/bin/bash
# round 2, call 7: blocked band-major companion (slices of 32 blocks of 8 entries): parity + timings + transpose geometries
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded_capi_gpu.py -m gpu -x -q -k "band_companion or sharded" > gpurun_out/pytest_gpu7.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu7.log
OUT=gpurun_out/opbench7.jsonl; : > $OUT
run() { local tag=$1; shift; local wl=$1; shift; local ops=$1; shift
  env "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops $ops --reps 10 --bmc 1 --tag $tag >> $OUT 2>> gpurun_out/opbench7.err; }
run b8 C4 spmv_t,spmv
run b4 C4 spmv_t,spmv SB200_BS_BLOCK=4
run b8 C2 spmv_t,spmv
run b4 C2 spmv_t,spmv SB200_BS_BLOCK=4
run b8 C3 spmv_t,spmv
run b8 C1 spmv_t,spmv
trun() { local tag=$1; shift; local wl=$1; shift
  env "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops transpose --reps 6 --tag $tag >> $OUT 2>> gpurun_out/opbench7.err; }
trun place_256x4096 C3
trun place_128x2048 C3 SB200_TRANSPOSE_CFG=128x2048
trun place_256x2048 C3 SB200_TRANSPOSE_CFG=256x2048
trun place_256x2048_nopf C3 SB200_TRANSPOSE_CFG=256x2048 SB200_TRANSPOSE_NOPF=1
trun place_nopf C3 SB200_TRANSPOSE_NOPF=1
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
grep build $OUT | tail -6
python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/plain_ncu_target7.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bandsweep -s 2 -c 1 -o gpurun_out/prof_bandsweep5_c2 \
  python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/ncu_bandsweep5.log 2>&1
echo "ncu rc=$?"
