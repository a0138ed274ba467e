#!/bin/bash
# round 2, call 47: lazy row indices for per-call mirrors + the block cache tests; the whole GPU suite; bench e2e section
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_lazy_rows_gpu.py tests/test_block_cache_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu47a.log 2>&1
echo "new tests rc=$?"; tail -4 gpurun_out/pytest_gpu47a.log
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu47.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu47.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --no-transpose --no-products --no-cpu-baseline --no-parity > gpurun_out/bench_e2e_lazy.json 2> gpurun_out/bench_e2e_lazy.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_e2e_lazy.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("e2e", e["ms_per_step"], "one-op", e["one_op_per_upload"]["ms_per_step"], e["one_op_per_upload"]["h2d_bytes_per_step"], "pageable", e["pageable"]["ms_per_step"], "transpose", e["transpose"]["ms_per_step"])
PY
python __graft_entry__.py --smoke 2>&1 | tail -2
