#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','row_path','gpu_launches','clocks')})
print({k:(round(v['ms'],4), round(v['frac_of_measured'],3)) for k,v in d['per_op'].items()})
print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])
PY
tail -3 gpurun_out/bench.err
for wl in C2 C3 C4; do echo "== $wl all ops"; timeout 900 python tools/opbench.py --workload $wl --reps 5 --warmup 2 > gpurun_out/op_$wl.jsonl 2>gpurun_out/op_$wl.err; cut -c60-330 gpurun_out/op_$wl.jsonl; done
