#!/bin/bash
# round 2, call 13 (2 GPUs): weak-scaling step with fewer exchange CTAs
mkdir -p gpurun_out
export SB200_EXCHANGE_TIMEOUT_S=60
for c in 296 148 74 32; do
  SB200_XG_CTAS=$c timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$((c%10)) bench.py --gpus 2 --steps 50 --warmup 10 --no-products --no-parity > gpurun_out/bench_n2_xg$c.json 2> gpurun_out/bench_n2_xg$c.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n2_xg$c.json").read().strip().splitlines()[-1])
print("ctas", $c, d["value"], d["ms_per_step"], {k: round(v["ms"], 4) for k, v in d["per_op"].items()})
PY
done
