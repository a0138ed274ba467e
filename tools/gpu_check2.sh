#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-400
echo "== e2e breakdown"
timeout 300 python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0,'.')
from rcppsparse_b200 import DeviceMatrix, synth
spec = synth.config("C2")
D = DeviceMatrix.synth(spec)
i,p,x = D.download_columns()
hi,hp,hx = (torch.from_numpy(a).pin_memory() for a in (i,p,x))
for validate in (True, False):
    for rep in range(3):
        torch.cuda.synchronize(); t0=time.perf_counter()
        M = DeviceMatrix.from_host(hi,hp,hx,D.nrow,D.ncol,validate=validate)
        t1=time.perf_counter()
        a=M.col_sums(); t2=time.perf_counter()
        b=M.row_sums(); t3=time.perf_counter()
        M.close(); t4=time.perf_counter()
        print(f"validate={validate} create {1e3*(t1-t0):.2f} ms  col_sums(host) {1e3*(t2-t1):.2f}  row_sums(host) {1e3*(t3-t2):.2f}  destroy {1e3*(t4-t3):.2f}")
# raw H2D speed
d = torch.empty(hx.numel(), dtype=torch.float64, device='cuda')
torch.cuda.synchronize(); t0=time.perf_counter(); d.copy_(hx, non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
print(f"torch H2D 800MB pinned: {1e3*(t1-t0):.2f} ms = {0.8/(t1-t0):.1f} GB/s")
PY
