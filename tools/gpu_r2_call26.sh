#!/bin/bash
# round 2, call 26: full GPU suite + driver-shaped bench with the two-split transpose in the product
mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu26.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu26.log
timeout -k 10 1500 python bench.py > gpurun_out/bench_n1_e.json 2> gpurun_out/bench_n1_e.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1_e.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"))
for k, v in d.get("roofline_by_op", {}).items():
    print("  ", k, round(v["ms"], 4), round(v["frac"], 3))
print("  layouts", d.get("row_companion"), d.get("product_layouts"))
PY
