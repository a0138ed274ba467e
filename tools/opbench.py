#!/usr/bin/env python
"""Per-op device timing on one GPU (tuning / ncu target; bench.py is the contract benchmark).

    python tools/opbench.py --workload C2 --ops colSums,rowSums --reps 20 [--scale 0.1]

Prints one JSON line per op: best/median CUDA-event ms, nnz/s, achieved GB/s (algorithmic bytes of
SURVEY.md 8d), fraction of the measured HBM peak.  Inputs are generated on the device.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ABI = {"colSums": "col_sums", "rowSums": "row_sums", "colMeans": "col_means", "rowMeans": "row_means", "spmv": "spmv",
       "spmv_t": "spmv_t", "transpose": "transpose"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2", help="C1..C4, or uniform:NROW:NCOL:DENSITY:SEED / powerlaw:NROW:NCOL:MEAN:SEED")
    ap.add_argument("--flush-l2", action="store_true", help="write a 512 MB buffer between repetitions (inputs smaller than L2)")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--ops", default="colSums,rowSums,colMeans,rowMeans,spmv,spmv_t,transpose")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tag", default="")
    ap.add_argument("--companion", type=int, default=0, help="1: build the row-ordered copy first, -1: never build it")
    ap.add_argument("--bmc", type=int, default=0, help="1: build both band-major companions first, -1: never build them")
    a = ap.parse_args()
    import torch

    from rcppsparse_b200 import DeviceMatrix, synth

    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    if ":" in a.workload:
        kind, nrow, ncol, par, seed = a.workload.split(":")
        if kind == "uniform":
            spec = synth.uniform_spec(int(nrow), int(ncol), float(par), int(seed), name=a.workload)
        elif kind == "powerlaw":
            spec = synth.powerlaw_spec(int(nrow), int(ncol), float(par), int(seed), name=a.workload)
        else:
            raise SystemExit(f"unknown workload kind {kind}")
    else:
        spec = synth.config(a.workload, a.scale)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda") if a.flush_l2 else None
    D = DeviceMatrix.synth(spec)
    D.set_stream(torch.cuda.current_stream().cuda_stream)
    if a.companion:
        D.row_companion(a.companion)
    if a.bmc:
        import time as _t
        for which in (0, 1):
            if a.bmc < 0 and which == 1 and a.companion >= 0:
                continue
            t0 = _t.perf_counter()
            D.band_companion(which, a.bmc)
            torch.cuda.synchronize()
            if a.bmc > 0:
                print(json.dumps({"tag": a.tag, "workload": spec.name, "build": f"band_companion({which})",
                                  "ms": round((_t.perf_counter() - t0) * 1e3, 2)}), flush=True)
    dev = torch.device("cuda", 0)
    out_c = torch.empty(max(D.ncol, 1), dtype=torch.float64, device=dev)
    out_r = torch.empty(max(D.nrow, 1), dtype=torch.float64, device=dev)
    v_c = torch.empty(max(D.ncol, 1), dtype=torch.float64, device=dev)
    v_r = torch.empty(max(D.nrow, 1), dtype=torch.float64, device=dev)
    D.synth_vector_dev(spec.seed, 0, D.ncol, v_c)
    D.synth_vector_dev(spec.seed + 7, 0, D.nrow, v_r)
    keep = []

    def run(op):
        if op == "colSums":
            D.col_sums_dev(out_c)
        elif op == "colMeans":
            D.col_sums_dev(out_c, float(D.nrow))
        elif op == "rowSums":
            D.row_sums_dev(out_r)
        elif op == "rowMeans":
            D.row_sums_dev(out_r, float(D.ncol))
        elif op == "spmv":
            D.spmv_dev(v_c, out_r)
        elif op == "spmv_t":
            D.spmv_t_dev(v_r, out_c)
        elif op == "transpose":
            keep.clear()
            keep.append(D.transpose_dev())

    for op in a.ops.split(","):
        for _ in range(a.warmup):
            run(op)
        torch.cuda.synchronize()
        ms = []
        for _ in range(a.reps):
            if flush is not None:
                flush.fill_(0.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(op)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        ab = D.algorithmic_bytes(ABI[op] + ("_companion" if op in ("rowSums", "rowMeans") and D.row_path() == "row-companion" else ""))
        best, med = float(np.min(ms)), float(np.median(ms))
        print(json.dumps({"tag": a.tag, "cfg": os.environ.get("SB200_SWEEP_CFG", ""), "workload": spec.name, "op": op, "row_path": D.row_path(),
                          "nnz": D.nnz, "ms_best": round(best, 4), "ms_median": round(med, 4),
                          "Gnnz_per_s": round(D.nnz / med / 1e6, 2), "GBps": round(ab / med / 1e6, 1),
                          "frac_measured": round(ab / med / 1e6 / peak, 3)}), flush=True)
    keep.clear()


if __name__ == "__main__":
    main()
