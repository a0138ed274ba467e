#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== all ops C2"; SB200_TRACE=1 timeout 600 python tools/opbench.py --workload C2 --reps 10 --tag v2 > gpurun_out/op_c2.jsonl 2> gpurun_out/op_c2.err; cat gpurun_out/op_c2.jsonl; tail -4 gpurun_out/op_c2.err
echo "== sweep cfg"
for cfg in 256,2,1 256,2,2 256,3,1 256,4,1 256,2,3 256,3,2 256,2,4; do
  SB200_SWEEP_CFG=$cfg timeout 300 python tools/opbench.py --workload C2 --ops colSums,spmv_t --reps 10 --tag cfg 2>> gpurun_out/op_cfg.err | tee -a gpurun_out/op_cfg.jsonl
done
echo "== C3"; SB200_TRACE=1 timeout 900 python tools/opbench.py --workload C3 --ops transpose,colSums,spmv_t --reps 3 --warmup 1 --tag c3 > gpurun_out/op_c3.jsonl 2> gpurun_out/op_c3.err; cat gpurun_out/op_c3.jsonl; tail -5 gpurun_out/op_c3.err
echo "== C4"; timeout 900 python tools/opbench.py --workload C4 --ops colSums,spmv_t --reps 3 --warmup 1 --tag c4 > gpurun_out/op_c4.jsonl 2> gpurun_out/op_c4.err; cat gpurun_out/op_c4.jsonl; tail -3 gpurun_out/op_c4.err
