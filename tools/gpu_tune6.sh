#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== C2 sweeps"; timeout 600 python tools/opbench.py --workload C2 --ops colSums,colMeans,spmv_t --reps 10 --tag v6 2>&1 | tail -3
for cfg in 256,2,1 256,3,1 256,4,1; do SB200_SWEEP_CFG=$cfg timeout 300 python tools/opbench.py --workload C2 --ops colSums,spmv_t --reps 10 --tag cfg 2>&1 | tail -2; done
echo "== C3/C4 colSums"; timeout 600 python tools/opbench.py --workload C3 --ops colSums,spmv_t --reps 3 --warmup 1 --tag c3 2>&1 | tail -2
timeout 600 python tools/opbench.py --workload C4 --ops colSums,spmv_t --reps 3 --warmup 1 --tag c4 2>&1 | tail -2
echo "== ncu"
timeout 300 python tools/opbench.py --workload C2 --ops rowSums,colSums,spmv_t --reps 1 --warmup 1 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'band_scatter_kernel|sweep_kernel' -c 6 -o gpurun_out/prof_scatter_v1 -f python tools/opbench.py --workload C2 --ops rowSums,colSums,spmv_t --reps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
