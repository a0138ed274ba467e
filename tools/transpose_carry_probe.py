#!/usr/bin/env python
"""C3 transpose (chunk-sort placement, bitmap ranks): full rounds with carry-over against whole windows, for several
band counts.  Every configuration's result is compared with the first one's through two sweeps over the transposed
matrix (column sums and a product with a random vector: bit-equal only if i' and x' are the same arrays).

    python tools/transpose_carry_probe.py [--workload C3] [--scale 1.0] [--reps 6]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--configs", default="0:0,1:0,1:148,1:200,1:110,0:148")
    a = ap.parse_args()
    import torch

    from rcppsparse_b200 import DeviceMatrix, synth

    spec = synth.config(a.workload, a.scale)
    D = DeviceMatrix.synth(spec)
    v = synth.dense_vector(7, spec.ncol)  # T has ncol(A) rows
    ref = None
    for cfg in a.configs.split(","):
        parts = cfg.split(":")
        carry, bands = parts[0], parts[1]
        geom = parts[2] if len(parts) > 2 else ""
        if geom:
            os.environ["SB200_TRANSPOSE_CFG"] = geom
        else:
            os.environ.pop("SB200_TRANSPOSE_CFG", None)
        os.environ["SB200_TRANSPOSE_CARRY"] = carry
        if int(bands) > 0:
            os.environ["SB200_TRANSPOSE_BANDS"] = bands
        else:
            os.environ.pop("SB200_TRANSPOSE_BANDS", None)
        T = D.transpose_dev()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
        torch.cuda.synchronize()
        for k in range(a.reps):
            ev[k].record()
            D.transpose_into(T)
        ev[a.reps].record()
        torch.cuda.synchronize()
        ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(a.reps)]
        sig = (T.col_sums(), T.spmv_t(v))
        same = None
        if ref is None:
            ref = sig
        else:
            same = bool(np.array_equal(ref[0].view(np.uint64), sig[0].view(np.uint64)) and
                        np.array_equal(ref[1].view(np.uint64), sig[1].view(np.uint64)))
        T.close()
        print(json.dumps({"carry": int(carry), "bands": int(bands), "cfg": geom or "default (256x1024 from 8e6 entries, else 256x2048)", "ms_median": float(np.median(ms)), "ms_min": float(min(ms)),
                          "same_as_first": same, "nnz": D.nnz}), flush=True)


if __name__ == "__main__":
    main()
