#!/usr/bin/env python
"""Wall-clock breakdown of a one-shot host-buffer transpose (create from pageable arrays + sb200_transpose into
pageable host vectors + destroy), the `e2e.transpose` figure of bench.py.  SB200_TRACE=1 adds the library's own phases.

    python tools/e2e_transpose_probe.py [--workload C2] [--reps 4]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    import torch

    from rcppsparse_b200 import DeviceMatrix, synth
    from rcppsparse_b200 import _lib as L
    from rcppsparse_b200.matrix import _ptr

    spec = synth.config(a.workload, a.scale)
    D = DeviceMatrix.synth(spec)
    i, p, x = D.download_columns()
    nnz = D.nnz
    nrow, ncol = D.nrow, D.ncol
    D.close()
    tp = np.empty(nrow + 1, np.int32)
    ti = np.empty(nnz, np.int32)
    tx = np.empty(nnz, np.float64)
    for rep in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        M = DeviceMatrix.from_host(i, p, x, nrow, ncol, device=0, validate=True)
        t1 = time.perf_counter()
        L.check(L.lib().sb200_transpose(M._h, _ptr(tp), _ptr(ti), _ptr(tx)))
        t2 = time.perf_counter()
        M.close()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"rep {rep}: total {(t3 - t0) * 1e3:.1f} ms  create {(t1 - t0) * 1e3:.1f}  transpose-to-host {(t2 - t1) * 1e3:.1f}  "
              f"destroy {(t3 - t2) * 1e3:.1f}", flush=True)


if __name__ == "__main__":
    main()
