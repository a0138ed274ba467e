import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
