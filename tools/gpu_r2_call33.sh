#!/bin/bash
# round 2, call 33: path selection by mean run length: parity (all transpose tests + full-size), default path on the
# crossover shapes
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu33.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu33.log
OUT=gpurun_out/opbench33.jsonl; : > $OUT; : > gpurun_out/opbench33.err
for wl in uniform:100000:100000:0.01:5 uniform:30000:1000000:0.02:8 uniform:30000:30000:0.01:3 uniform:100000:50000:0.002:4 C1 C2 C3; do
  SB200_TRACE=1 timeout -k 10 300 python tools/opbench.py --workload $wl --ops transpose --reps 5 --tag default >> $OUT 2>> gpurun_out/opbench33.err
done
for wl in uniform:30000:30000:0.01:3 uniform:100000:50000:0.002:4; do
  for path in place split; do
    SB200_TRANSPOSE_PATH=$path timeout -k 10 300 python tools/opbench.py --workload $wl --ops transpose --reps 5 --tag $path >> $OUT 2>> gpurun_out/opbench33.err
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/opbench33.jsonl"):
    d = json.loads(l)
    if "op" in d: print(d["tag"], d["workload"][:40], d["nnz"], d["ms_median"], d["frac_measured"])
PY
grep trace gpurun_out/opbench33.err | grep "plan built" | cut -c1-120
