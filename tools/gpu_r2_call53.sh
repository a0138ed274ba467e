#!/bin/bash
# round 2, call 53: one ncu capture (full set, source) of the chunk-sort placement with carried windows, C3 at 0.15 scale
mkdir -p gpurun_out
python tools/opbench.py --workload C3 --scale 0.15 --ops transpose --reps 3 > gpurun_out/plain_ncu_target53.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bitrank -s 2 -c 1 -f -o gpurun_out/prof_bitrank_carry_c3s \
  python tools/opbench.py --workload C3 --scale 0.15 --ops transpose --reps 3 > gpurun_out/ncu_bitrank53.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/plain_ncu_target53.log | cut -c1-250; ls -la gpurun_out/prof_bitrank_carry_c3s.ncu-rep
