#!/usr/bin/env python
"""Wall-clock breakdown of one host-buffer step (upload + four sweeps + result read-back), C2 by default.

    python tools/e2e_phases.py [--workload C2] [--reps 5]
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
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch

    from rcppsparse_b200 import DeviceMatrix, synth

    spec = synth.config(a.workload, a.scale)
    D = DeviceMatrix.synth(spec)
    i, p, x = D.download_columns()
    hi, hp, hx = (torch.from_numpy(t).pin_memory() for t in (i, p, x))
    nbytes = hi.numel() * 4 + hp.numel() * 4 + hx.numel() * 8
    # raw pinned copy rate of this box
    di, dx = torch.empty_like(hi, device="cuda"), torch.empty_like(hx, device="cuda")
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        di.copy_(hi, non_blocking=True)
        dx.copy_(hx, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"raw pinned H2D: {nbytes / 1e6:.1f} MB in {dt * 1e3:.2f} ms = {nbytes / dt / 1e9:.1f} GB/s")
    del di, dx
    for rep in range(a.reps):
        ph = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        M = DeviceMatrix.from_host(hi, hp, hx, D.nrow, D.ncol, device=0, validate=True)
        ph["create"] = time.perf_counter() - t0
        for name in ("col_sums", "row_sums", "col_means", "row_means"):
            t1 = time.perf_counter()
            getattr(M, name)()
            ph[name] = time.perf_counter() - t1
        t1 = time.perf_counter()
        M.close()
        ph["destroy"] = time.perf_counter() - t1
        tot = time.perf_counter() - t0
        print(f"rep {rep}: total {tot * 1e3:.2f} ms  " + "  ".join(f"{k} {v * 1e3:.2f}" for k, v in ph.items()) +
              f"  -> {4 * D.nnz / tot / 1e9:.2f} Gnnz/s")


if __name__ == "__main__":
    main()
