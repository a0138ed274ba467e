#!/bin/bash
# round 2, call 38 (2 GPUs): sharded transpose behind the C ABI; the refactored push kernel of the cross-process exchange
mkdir -p gpurun_out
export SB200_EXCHANGE_TIMEOUT_S=60
timeout -k 10 900 python -m pytest tests/test_sharded_capi_gpu.py tests/test_exchange_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu38.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu38.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/shard_check.py > gpurun_out/shard_check_n2c.log 2>&1
echo "shard_check rc=$?"; tail -2 gpurun_out/shard_check_n2c.log | cut -c1-200
python - <<'PY'
import time, numpy as np
from rcppsparse_b200 import ShardedHostMatrix, synth
spec = synth.config("C2")
i, p, x = synth.generate_host(spec)
for g in (1, 2):
    with ShardedHostMatrix(i, p, x, spec.nrow, spec.ncol, g) as S:
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); S.transpose_host(); ts.append(round((time.perf_counter() - t0) * 1e3, 1))
        print("C2 sharded transpose to host, gpus", g, "ms", ts)
PY
