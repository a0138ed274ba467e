#!/bin/bash
# round 2, call 42: pageable staging threads (4/8/12/16) on the one-shot transpose; rebuild time of the row-ordered copy under trace
mkdir -p gpurun_out
for t in 4 8 12 16; do
  SB200_COPY_THREADS=$t timeout -k 10 300 python tools/e2e_transpose_probe.py --reps 6 > gpurun_out/copy_threads_$t.log 2>&1
  echo "threads $t:"; grep "^rep" gpurun_out/copy_threads_$t.log | tail -3
done
SB200_TRACE=1 python - > gpurun_out/row_copy_rebuild_trace.log 2>&1 <<'PY'
import time, torch
from rcppsparse_b200 import DeviceMatrix, synth
D = DeviceMatrix.synth(synth.config("C2"))
D.row_companion(1)
for k in range(5):
    D.row_companion(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D.row_companion(1)
    torch.cuda.synchronize()
    print(f"rebuild {k}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
PY
tail -25 gpurun_out/row_copy_rebuild_trace.log
