import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from rcppsparse_b200 import DeviceMatrix, synth
spec = synth.config("C2")
D = DeviceMatrix.synth(spec)
i, p, x = D.download_columns()   # pageable numpy arrays, like R-owned vectors
for pin in (False, True, False, True):
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        M = DeviceMatrix.from_host(i, p, x, D.nrow, D.ncol, pin=pin)
        ts.append((time.perf_counter() - t0) * 1e3); M.close()
    print("create from pageable memory, pin =", pin, [round(t, 1) for t in ts], "ms")
ti = np.empty(D.nnz, np.int32); tp = np.empty(D.nrow + 1, np.int32); tx = np.empty(D.nnz)
for _ in range(3):
    t0 = time.perf_counter(); r = D.transpose_host(); print("transpose to pageable host", round((time.perf_counter() - t0) * 1e3, 1), "ms")
