#!/usr/bin/env python
"""What does the FIRST transpose of a matrix cost (plan + scratch + kernels) in a warm process?  Three matrices of the
same shape, different seeds, one after the other; wall clock around sb200_transpose_dev + synchronize."""
import argparse
import dataclasses
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch

    from rcppsparse_b200 import DeviceMatrix, synth

    base = synth.config(a.workload)
    out = []
    for k in range(3):
        spec = dataclasses.replace(base, seed=base.seed + 1000 * k)
        D = DeviceMatrix.synth(spec)
        torch.cuda.synchronize()
        ms = []
        for call in range(3):
            t0 = time.perf_counter()
            T = D.transpose_dev()
            torch.cuda.synchronize()
            ms.append(round((time.perf_counter() - t0) * 1e3, 2))
            T.close()
        D.close()
        out.append(ms)
    print(json.dumps({"tag": a.tag, "workload": base.name, "path": os.environ.get("SB200_TRANSPOSE_PATH", "default"),
                      "ms_calls_1_2_3_per_matrix": out}))


if __name__ == "__main__":
    main()
