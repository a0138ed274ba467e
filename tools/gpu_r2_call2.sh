#!/bin/bash
# round 2, call 2: band sweep ring / prefetch variants, then one ncu capture of the kernel at C2
mkdir -p gpurun_out
OUT=gpurun_out/opbench_bs_tune.jsonl
: > $OUT
run() {  # tag, env..., workload
  local tag=$1; shift; local wl=$1; shift
  env "$@" timeout -k 10 300 python tools/opbench.py --workload $wl --ops spmv_t,spmv --reps 10 --bmc 1 --tag $tag >> $OUT 2>> gpurun_out/opbench_bs_tune.err
}
for wl in C4 C2; do
  run ahead0_1024x2 $wl SB200_BS_AHEAD=0
  run ahead12k_1024x2 $wl SB200_BS_AHEAD=12288
  run ahead32k_1024x2 $wl SB200_BS_AHEAD=32768
  run ahead12k_640x3 $wl SB200_BS_AHEAD=12288 SB200_BS_CFG=640,3
  run ahead12k_512x4 $wl SB200_BS_AHEAD=12288 SB200_BS_CFG=512,4
  run ahead0_512x4 $wl SB200_BS_AHEAD=0 SB200_BS_CFG=512,4
done
run ahead12k_C3 C3 SB200_BS_AHEAD=12288
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/plain_ncu_target.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bandsweep -s 2 -c 2 -o gpurun_out/prof_bandsweep_c2 \
  python tools/opbench.py --workload C2 --ops spmv_t --reps 3 --bmc 1 > gpurun_out/ncu_bandsweep.log 2>&1
echo "ncu rc=$?"
