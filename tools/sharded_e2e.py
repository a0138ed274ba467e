#!/usr/bin/env python
"""End-to-end step of ONE process over 1..n GPUs through sb200_sharded_* (the path the drop-in header takes with
SB200_GPUS=n): every step cuts the host dgCMatrix (pageable arrays, as R owns them) into column blocks, uploads block k
to GPU k over its own PCIe link, runs colSums + rowSums + colMeans + rowMeans into host vectors and destroys the
mirrors.  Also the one-shot transpose into host arrays.  Wall clock, median of the timed steps.

    python tools/sharded_e2e.py [--workload C2] [--steps 5] [--out gpurun_out/sharded_e2e.json]
"""
import argparse
import json
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch

    from rcppsparse_b200 import ShardedHostMatrix, synth

    spec = synth.config(a.workload, a.scale)
    i, p, x = synth.generate_host(spec)
    nnz = int(x.shape[0])
    ndev = torch.cuda.device_count()
    res = {"workload": a.workload, "nnz": nnz, "devices": ndev, "steps": a.steps, "by_gpus": {}}
    ref = None
    for g in (1, 2, 4, 8):
        if g > ndev:
            break

        def step():
            with ShardedHostMatrix(i, p, x, spec.nrow, spec.ncol, g) as S:
                return S.col_sums(), S.row_sums(), S.col_means(), S.row_means()

        def tstep():
            with ShardedHostMatrix(i, p, x, spec.nrow, spec.ncol, g) as S:
                return S.transpose_host()

        out = {}
        for name, fn, ops in (("four_sums", step, 4), ("transpose", tstep, 1)):
            r = fn()
            ts = []
            for _ in range(a.steps):
                t0 = time.perf_counter()
                r = fn()
                ts.append(time.perf_counter() - t0)
            dt = float(np.median(ts))
            out[name] = {"ms_per_step": dt * 1e3, "nnz_per_s": ops * nnz / dt, "ms_all": [round(t * 1e3, 2) for t in ts]}
            if name == "four_sums":
                if ref is None:
                    ref = r
                else:  # a different cut moves tile boundaries and the order of the partial sums: inside the parity bar, not bit-equal
                    for k in range(4):
                        assert np.allclose(ref[k], r[k], rtol=1e-10, atol=1e-9), f"result {k} differs between 1 and {g} GPUs"
            del r
        res["by_gpus"][str(g)] = out
        print(f"gpus {g}: four sums {out['four_sums']['ms_per_step']:.1f} ms/step = {out['four_sums']['nnz_per_s'] / 1e9:.1f} Gnnz/s; "
              f"one-shot transpose {out['transpose']['ms_per_step']:.1f} ms", flush=True)
    line = json.dumps(res)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")
    print(line)


if __name__ == "__main__":
    main()
