#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== all ops C2"; timeout 600 python tools/opbench.py --workload C2 --reps 10 --tag base > gpurun_out/op_c2.jsonl 2> gpurun_out/op_c2.err; cat gpurun_out/op_c2.jsonl; tail -3 gpurun_out/op_c2.err
echo "== sweep cfg"
for cfg in 512,2,1 512,2,2 512,3,1 512,4,1 512,3,2; do
  SB200_SWEEP_CFG=$cfg timeout 300 python tools/opbench.py --workload C2 --ops colSums,spmv_t,spmv --reps 10 --tag cfg 2>> gpurun_out/op_cfg.err | tee -a gpurun_out/op_cfg.jsonl
done
echo "== C3 transpose + sweeps"; timeout 900 python tools/opbench.py --workload C3 --ops transpose,colSums,rowSums,spmv,spmv_t --reps 3 --warmup 1 --tag c3 > gpurun_out/op_c3.jsonl 2> gpurun_out/op_c3.err; cat gpurun_out/op_c3.jsonl; tail -3 gpurun_out/op_c3.err
echo "== C4 spmv"; timeout 900 python tools/opbench.py --workload C4 --ops spmv,spmv_t,colSums,rowSums --reps 3 --warmup 1 --tag c4 > gpurun_out/op_c4.jsonl 2> gpurun_out/op_c4.err; cat gpurun_out/op_c4.jsonl; tail -3 gpurun_out/op_c4.err
echo "== ncu"
timeout 300 python tools/opbench.py --workload C2 --ops colSums,rowSums,spmv_t --reps 1 --warmup 1 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sweep_kernel|rowsum_stream' -c 6 -o gpurun_out/prof_sweep_v1 -f python tools/opbench.py --workload C2 --ops colSums,rowSums,spmv_t --reps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu.log
