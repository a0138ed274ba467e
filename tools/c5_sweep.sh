#!/bin/bash
# SURVEY.md 8(d) "C5": N in {1e6 .. 2e9} x {uniform d=1e-3 over 1M rows, power-law columns (mean 1000) over 2^20 rows}
# x all seven ops on one GPU; L2 flushed between repetitions (the small points fit the 126 MB L2).
set +e
mkdir -p gpurun_out
: > gpurun_out/c5_sweep.jsonl
seed=1005
for n in 1000 10000 100000 1000000 2000000; do
  for kind in "uniform:1000000:$n:0.001" "powerlaw:1048576:$n:1000.0"; do
    reps=10; [ $n -ge 1000000 ] && reps=4
    timeout 600 python tools/opbench.py --workload "$kind:$seed" --reps $reps --warmup 2 --flush-l2 --tag c5 >> gpurun_out/c5_sweep.jsonl 2>> gpurun_out/c5_sweep.err
    seed=$((seed+1))
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/c5_sweep.jsonl'):
    d=json.loads(l); print(f"{d['workload']:34s} {d['op']:9s} nnz {d['nnz']:>11d} {d['ms_median']:>10.4f} ms {d['GBps']:>8.1f} GB/s frac {d['frac_measured']:.3f} ({d['row_path']})")
PY
