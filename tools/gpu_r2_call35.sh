#!/bin/bash
# round 2, call 35: final parity run after the last selection change (all GPU tests), small-matrix transposes
mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu35.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu35.log
for wl in uniform:1000000:1000:0.001:1005 uniform:1000000:4000:0.001:1006 uniform:1000000:10000:0.001:1007; do
  timeout -k 10 300 python tools/opbench.py --workload $wl --ops transpose --reps 8 --flush-l2 --tag default 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], d['nnz'], d['ms_median'])"
done
