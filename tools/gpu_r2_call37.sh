#!/bin/bash
# round 2, call 37: transpose_into (test) and the bench line with it
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "transpose" > gpurun_out/pytest_gpu37.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu37.log
bash tools/gpu_r2_call36.sh
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1_g.json").read().strip().splitlines()[-1])
for k in ("transpose@C3", "transpose@C2", "transpose@C4"):
    v = d["roofline_by_op"][k]; print(k, v["ms_per_launch"], v["ms_min"], v["allocating_form_ms"])
PY
