#!/bin/bash
# round 2, call 31: ncu on the final two-split transpose (C2), new 5M-row test
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "five_million or two_stream" > gpurun_out/pytest_gpu31.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu31.log
python tools/opbench.py --workload C2 --ops transpose --reps 3 > gpurun_out/plain_ncu_target31.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:split_kernel -s 2 -c 2 -o gpurun_out/prof_split5_c2 \
  python tools/opbench.py --workload C2 --ops transpose --reps 3 > gpurun_out/ncu_split5.log 2>&1
echo "ncu rc=$?"
