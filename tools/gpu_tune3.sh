#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (transpose)"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "transpose or smoke or round_trip" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== C2 transpose"; SB200_TRACE=1 timeout 600 python tools/opbench.py --workload C2 --ops transpose --reps 3 --tag v3 2>&1 | tail -3
echo "== C3 transpose"; SB200_TRACE=1 timeout 900 python tools/opbench.py --workload C3 --ops transpose --reps 3 --warmup 1 --tag c3 2>&1 | tail -3
for kw in 4 8; do echo "== C3 kw=$kw"; SB200_TRANSPOSE_KW=$kw SB200_TRACE=1 timeout 900 python tools/opbench.py --workload C3 --ops transpose --reps 2 --warmup 1 --tag c3kw$kw 2>&1 | tail -2; done
echo "== C3 bands=148"; SB200_TRANSPOSE_BANDS=148 SB200_TRACE=1 timeout 900 python tools/opbench.py --workload C3 --ops transpose --reps 2 --warmup 1 --tag c3b148 2>&1 | tail -2
echo "== ncu"
timeout 300 python tools/opbench.py --workload C3 --scale 0.1 --ops transpose,colSums --reps 1 --warmup 1 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'band_ptr_kernel|transpose_band_kernel|sweep_kernel|row_hist' -c 8 -o gpurun_out/prof_transpose_v2 -f python tools/opbench.py --workload C3 --scale 0.1 --ops transpose,colSums --reps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
