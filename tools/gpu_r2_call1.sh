#!/bin/bash
# round 2, call 1: parity of the new band companion, per-op timings, first full bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/box.txt 2>&1
nproc >> gpurun_out/box.txt; free -g | head -2 >> gpurun_out/box.txt
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for W in C2 C3 C4; do
  timeout -k 10 600 python tools/opbench.py --workload $W --ops spmv_t,spmv --reps 10 --bmc 1 --tag bmc >> gpurun_out/opbench_bmc.jsonl 2>> gpurun_out/opbench_bmc.err
  echo "opbench $W rc=$?"
done
for L in 4 8 32; do
  SB200_BS_LANES=$L timeout -k 10 600 python tools/opbench.py --workload C4 --ops spmv_t,spmv --reps 10 --bmc 1 --tag lanes$L >> gpurun_out/opbench_bmc.jsonl 2>> gpurun_out/opbench_bmc.err
done
tail -20 gpurun_out/opbench_bmc.jsonl
timeout -k 10 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_n1.err
