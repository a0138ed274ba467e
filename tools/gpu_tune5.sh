#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== C2"; SB200_TRACE=1 timeout 600 python tools/opbench.py --workload C2 --ops rowSums,spmv,transpose --reps 5 --tag v5 2>&1 | grep -v "band plan" | tail -5
for sp in 2 8; do echo "== C2 scatter splits=$sp"; SB200_SCATTER_SPLITS=$sp timeout 600 python tools/opbench.py --workload C2 --ops rowSums --reps 5 --tag sp$sp 2>&1 | tail -1; done
echo "== C3"; SB200_TRACE=1 timeout 900 python tools/opbench.py --workload C3 --ops transpose,rowSums,spmv --reps 3 --warmup 1 --tag c3 2>&1 | tail -6
echo "== C4"; timeout 900 python tools/opbench.py --workload C4 --ops rowSums,spmv --reps 3 --warmup 1 --tag c4 2>&1 | tail -3
