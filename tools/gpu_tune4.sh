#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== C2"; SB200_TRACE=1 timeout 600 python tools/opbench.py --workload C2 --ops rowSums,rowMeans,spmv,transpose --reps 5 --tag v4 2>&1 | tail -9
for sp in 1 2 8; do echo "== C2 scatter splits=$sp"; SB200_SCATTER_SPLITS=$sp timeout 600 python tools/opbench.py --workload C2 --ops rowSums,spmv --reps 5 --tag sp$sp 2>&1 | tail -2; done
echo "== C2 bands=296"; SB200_BANDS=296 SB200_SCATTER_SPLITS=2 timeout 600 python tools/opbench.py --workload C2 --ops rowSums,spmv --reps 5 --tag b296 2>&1 | tail -2
echo "== C3"; SB200_TRACE=1 timeout 900 python tools/opbench.py --workload C3 --ops transpose,rowSums,spmv --reps 3 --warmup 1 --tag c3 2>&1 | tail -8
echo "== C3 transpose S=4"; SB200_TRANSPOSE_SPLITS=4 SB200_TRACE=1 timeout 900 python tools/opbench.py --workload C3 --ops transpose --reps 2 --warmup 1 --tag c3s4 2>&1 | tail -3
echo "== C3 transpose bands=74 S=4"; SB200_TRANSPOSE_BANDS=74 SB200_TRANSPOSE_SPLITS=4 SB200_TRACE=1 timeout 900 python tools/opbench.py --workload C3 --ops transpose --reps 2 --warmup 1 --tag c3b74 2>&1 | tail -3
echo "== C4"; timeout 900 python tools/opbench.py --workload C4 --ops rowSums,spmv --reps 3 --warmup 1 --tag c4 2>&1 | tail -3
