#!/bin/bash
# One gpurun call: smoke, GPU parity tests, micro-benchmarks, bench line.  Each step has its own timeout
# so a hung kernel cannot hold the box.  Outputs land in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== microbench"; timeout 300 ./rcppsparse_b200/microbench > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?"; cat gpurun_out/microbench.jsonl; tail -3 gpurun_out/microbench.err
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
