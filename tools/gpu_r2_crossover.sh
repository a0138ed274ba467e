#!/bin/bash
# round 2: where does the two-split transpose overtake the chunk sort?  matrices between C3 (short-wide) and C2 (tall)
mkdir -p gpurun_out
OUT=gpurun_out/opbench_crossover.jsonl; : > $OUT; : > gpurun_out/opbench_crossover.err
for wl in uniform:100000:100000:0.01:5 uniform:60000:400000:0.005:6 uniform:200000:100000:0.004:7 uniform:30000:1000000:0.02:8 powerlaw:100000:200000:800:9; do
  for path in place split; do
    SB200_TRANSPOSE_PATH=$path timeout -k 10 300 python tools/opbench.py --workload $wl --ops transpose --reps 5 --tag $path >> $OUT 2>> gpurun_out/opbench_crossover.err
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/opbench_crossover.jsonl"):
    d = json.loads(l)
    if "op" in d: print(d["tag"], d["workload"], d["nnz"], d["ms_median"], d["frac_measured"])
PY
