#!/bin/bash
# round 2, call 36: bench with transpose@C2 / transpose@C4 in the line
mkdir -p gpurun_out
timeout -k 10 1500 python bench.py > gpurun_out/bench_n1_g.json 2> gpurun_out/bench_n1_g.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n1_g.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1_g.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), d.get("gpu_launches"), d.get("wall_s_timed_region"))
for k, v in d.get("roofline_by_op", {}).items():
    print("  ", k, round(v["ms_per_launch"], 4), round(v["frac"], 3), v.get("first_call_ms"), v.get("check"))
print({k: (v.get("error") if isinstance(v, dict) else v) for k, v in d.get("sections", {}).items()})
PY
