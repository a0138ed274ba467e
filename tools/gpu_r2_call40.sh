#!/bin/bash
# round 2, call 40: is the ~440 ms stall of the one-shot transpose reproducible? (12 reps of the probe; the bench's own e2e section under trace)
mkdir -p gpurun_out
SB200_TRACE=1 timeout -k 10 600 python tools/e2e_transpose_probe.py --reps 12 > gpurun_out/e2e_transpose_probe.log 2>&1
echo "probe rc=$?"; grep "^rep\|took" gpurun_out/e2e_transpose_probe.log | tail -40
SB200_TRACE=1 timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-transpose --no-products --no-cpu-baseline --no-parity > gpurun_out/bench_e2e_trace.json 2> gpurun_out/bench_e2e_trace.err
echo "bench rc=$?"
grep -n "transpose to host\|two stream splits\|took\|split plan" gpurun_out/bench_e2e_trace.err | tail -60
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_e2e_trace.json").read().strip().splitlines()[-1])
print(json.dumps(d["e2e"], indent=1)[:1500])
PY
