#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in C2 C3; do echo "== $wl"; timeout 600 python tools/opbench.py --workload $wl --ops rowSums,rowMeans,spmv --reps 5 --warmup 2 2>&1 | tail -3 | cut -c110-330; done
echo "== C4 banded"; SB200_ROW_PLAN=1 timeout 600 python tools/opbench.py --workload C4 --ops rowSums,spmv --reps 3 --warmup 1 2>&1 | tail -2 | cut -c110-330
