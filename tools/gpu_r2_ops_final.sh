#!/bin/bash
# round 2, final per-op table: every op on every config, on the CSC arrays and on the cached layouts
mkdir -p gpurun_out
OUT=gpurun_out/opbench_final.jsonl; : > $OUT; : > gpurun_out/opbench_final.err
for wl in C1 C2 C3 C4; do
  timeout -k 10 600 python tools/opbench.py --workload $wl --reps 8 --companion -1 --bmc -1 --tag csc_arrays >> $OUT 2>> gpurun_out/opbench_final.err
  timeout -k 10 600 python tools/opbench.py --workload $wl --reps 8 --companion 1 --bmc 1 --ops rowSums,rowMeans,spmv,spmv_t --tag cached_layouts >> $OUT 2>> gpurun_out/opbench_final.err
done
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
