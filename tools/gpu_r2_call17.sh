#!/bin/bash
# round 2, call 17: ncu on the two stream-split passes (C2)
mkdir -p gpurun_out
python tools/opbench.py --workload C2 --ops transpose --reps 3 > gpurun_out/plain_ncu_target17.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:split_kernel -s 2 -c 2 -o gpurun_out/prof_split_c2 \
  python tools/opbench.py --workload C2 --ops transpose --reps 3 > gpurun_out/ncu_split.log 2>&1
echo "ncu rc=$?"
