#!/bin/bash
# round 2, call 8 (2 GPUs): parity subset, C3 products after the warp fold, bench at N=1 and N=2
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "band_companion or transpose" > gpurun_out/pytest_gpu8.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu8.log
OUT=gpurun_out/opbench8.jsonl; : > $OUT
for W in C3 C4 C2; do
  timeout -k 10 300 python tools/opbench.py --workload $W --ops spmv_t,spmv --reps 10 --bmc 1 --tag auto >> $OUT 2>> gpurun_out/opbench8.err
done
grep -v build $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], d['workload'][:2], d['op'], d['ms_median'], d['frac_measured'])"
timeout -k 10 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err
echo "bench n1 rc=$?"; tail -3 gpurun_out/bench_n1_b.err
timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_b.json 2> gpurun_out/bench_n2_b.err
echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_n2_b.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1_b.json", "gpurun_out/bench_n2_b.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"])
        for k, v in d["roofline_by_op"].items():
            print("   ", k, round(v["ms_per_launch"], 4), round(v["frac"], 3))
        print("    c4_strong", d.get("c4_strong"))
        print("    parity", d.get("parity", {}).get("worst_err_over_sum_abs_terms"))
        print("    sections", {k: (v.get("error") if isinstance(v, dict) else v) for k, v in d.get("sections", {}).items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
