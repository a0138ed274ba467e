#!/bin/bash
# round 2, final evidence: full GPU suite, the driver-shaped bench line, the ncu launch list of the same bench command
mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout -k 10 1500 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
echo "bench rc=$?"
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref_final.json
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/plain_bench_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench_final.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_final.log 2>&1
echo "ncu launches rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), d.get("gpu_launches"), d.get("clocks"))
for k, v in d.get("roofline_by_op", {}).items():
    print("  ", k, round(v["ms_per_launch"], 4), round(v["frac"], 3))
rc = d.get("row_companion", {})
print("  row copy build", rc.get("build_ms"), "rebuild", rc.get("rebuild_ms_pool_warm"))
PY
