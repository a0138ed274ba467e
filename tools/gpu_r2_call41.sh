#!/bin/bash
# round 2, call 41: block cache over the driver pool: one-shot transpose probe (12 reps), full GPU suite, bench e2e section
mkdir -p gpurun_out
SB200_TRACE=1 timeout -k 10 600 python tools/e2e_transpose_probe.py --reps 12 > gpurun_out/e2e_transpose_probe_cache.log 2>&1
echo "probe rc=$?"; grep "^rep\|took" gpurun_out/e2e_transpose_probe_cache.log | tail -30
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu41.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu41.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-transpose --no-products --no-cpu-baseline --no-parity > gpurun_out/bench_e2e_cache.json 2> gpurun_out/bench_e2e_cache.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_e2e_cache.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("e2e", e["ms_per_step"], "one-op", e["one_op_per_upload"]["ms_per_step"], "pageable", e["pageable"]["ms_per_step"], "transpose", e["transpose"]["ms_per_step"])
print("value", d["value"], d["row_companion"]["build_ms"], d["row_companion"]["rebuild_ms_pool_warm"])
PY
