#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the RcppSparse hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own serial CPU code

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): 1,000,000 x 100,000 uniform
rsparsematrix-style dgCMatrix, density 1e-3 (~1e8 stored entries, FP64 values, int32 indices),
synthetic, generated straight into HBM by the integer-exact recipe of rcppsparse_b200/synth.py.
One STEP = one pass of the four reductions the config names — colSums, rowSums, colMeans,
rowMeans (reference RcppSparse.h:131-156) — over the device-resident matrix.
`value` = stored entries swept per second over the whole job = ops * nnz / step time.

At N > 1 (torchrun, one rank per GPU) every rank owns a C2-sized column block of a
1M x (100k*N) matrix (weak scaling; columns are independent units); inside the timed step column
results are assembled and row results summed by the library's exchange kernels over NVLink peer
memory (--exchange nccl: torch.distributed all-gather / all-reduce instead), each op's exchange running
beside the next op's sweep, every result waited for before its step ends.

Row sums of a RESIDENT mirror: the library serves rowSums/rowMeans with its scatter kernels until a mirror has
been asked for them more than 8 times, then from a row-ordered copy it builds once (sparse_b200.h).  The
steady state measured here is the second regime; the first is timed before warm-up and reported in
`row_companion` (scatter time per call, one-off build time, and `value_before_row_copy`: the same step
with the row sums still on the scatter kernel).  --no-row-companion keeps the whole run in the first regime.

Extra keys beside the base contract: `per_op`, `roofline` (slowest op's kernel: algorithmic bytes per launch
from sb200_algorithmic_bytes, timed live with CUDA events on the launching stream, `traffic` from the
committed ncu capture), `row_path`, `row_companion`, `exchange` (N > 1), `cpu_baseline` (reference code on a
bounded column block, rank 0, N=1), `e2e` (same step through the host-buffer C ABI: upload of i/p/x from pinned
memory + four results read back into pinned buffers, every step; `e2e.pageable`: the same from ordinary host
arrays, what an R caller owns), `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS = ("colSums", "rowSums", "colMeans", "rowMeans")
ABI_OP = {"colSums": "col_sums", "rowSums": "row_sums", "colMeans": "col_means", "rowMeans": "row_means",
          "spmv": "spmv", "spmv_t": "spmv_t", "transpose": "transpose"}
METRIC = "nnz/s"
NOMINAL_HBM_GBS = 8000.0  # BASELINE.json metric: "% of 8 TB/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", help="C1|C2|C3|C4 (default C2 = BASELINE configs[1])")
    ap.add_argument("--scale", type=float, default=1.0, help="column-count scale of the workload (tests)")
    ap.add_argument("--ops", default=",".join(OPS), help="comma list; also spmv,spmv_t,transpose")
    ap.add_argument("--cpu-cols", type=int, default=25000, help="columns of the block timed on the CPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--exchange", default=None, choices=["p2p", "nccl"],
                    help="N > 1: the library's peer-memory exchange kernels (default) or NCCL collectives")
    ap.add_argument("--overlap", action="store_true",
                    help="N > 1 with --exchange nccl: leave each op's collective in flight under the next op's sweep "
                         "(measured slower: the NCCL kernel waits for SM resources behind the persistent sweep)")
    ap.add_argument("--no-row-companion", action="store_true",
                    help="keep rowSums/rowMeans on the scatter kernels (no row-ordered copy of the resident mirror)")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def value_before_row_copy(ops, nnz, per_op_ms, scatter_ms):
    """Throughput of the same step with rowSums/rowMeans still on the scatter kernels (a mirror's first 8 row-sum
    calls, and adopted device arrays): the timed ops' own times, with each row op replaced by the scatter time
    measured before the row-ordered copy was built."""
    total_ms = sum(scatter_ms if op in ("rowSums", "rowMeans") else per_op_ms[op] for op in ops)
    return len(ops) * nnz / (total_ms * 1e-3) if total_ms > 0 else None


def profile_traffic(kernel):
    """dram bytes per launch of `kernel` at the default workload from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    try:
        return json.load(open(path))["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t0 = self.t1 = None

    def start(self):
        """Started well before the timed region (nvidia-smi needs ~1 s to produce its first line)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first_sample(self, timeout=5.0):
        t_end = time.perf_counter() + timeout
        while self.proc and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.02)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= self.t1 + 0.05]
        where = "during the timed region"
        if not inside:  # region shorter than one sampling period: take the samples right around it
            inside = [r for (t, r) in self.rows if self.t0 is not None and abs(t - self.t0) < 0.5]
            where = "within 0.5 s of the timed region (region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "sampled": where}


# =====================================================================================================
# reference arm: the reference's own serial CPU code (oracle/_ref when it was built from
# /root/reference, else the C port of the same loops) on a bounded column block of the same workload
# =====================================================================================================
def cpu_block(spec, n_cols):
    from rcppsparse_b200 import synth

    n_cols = max(1, min(n_cols, spec.ncol))
    i, p, x = synth.generate_host(spec, 0, n_cols)
    return i, p, x, spec.nrow, n_cols


def time_cpu(ops, block, reps):
    """Best-of-reps wall time of each op on the block; returns (checker kind, {op: seconds}, nnz)."""
    from oracle import oracle
    from rcppsparse_b200 import synth

    chk = oracle.best()
    i, p, x, nrow, ncol = block
    v_c, v_r = synth.dense_vector(1, ncol), synth.dense_vector(2, nrow)
    out = {}
    for op in ops:
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            if op == "spmv":
                chk.spmv(i, p, x, nrow, ncol, v_c)
            elif op == "spmv_t":
                chk.spmv_t(i, p, x, nrow, ncol, v_r)
            else:
                getattr(chk, op)(i, p, x, nrow, ncol)
            best = min(best, time.perf_counter() - t0)
        out[op] = best
    return chk.kind, out, int(x.shape[0])


def run_reference(args, ops):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    from rcppsparse_b200 import synth

    spec = synth.config(args.workload, args.scale)
    # bounded sample: keep the whole --steps/--warmup run within about a minute of CPU work
    cols = max(1000, min(args.cpu_cols, int(args.cpu_cols * 100 / max(1, args.steps + args.warmup))))
    block = cpu_block(spec, cols)
    nnz = int(block[2].shape[0])
    for _ in range(max(1, min(args.warmup, 3))):
        time_cpu(ops, block, 1)
    kind = "port"
    step_s = []
    for _ in range(args.steps):
        kind, t, _ = time_cpu(ops, block, 1)
        step_s.append(sum(t.values()))
    step = float(np.median(step_s))
    value = len(ops) * nnz / step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(spec, args, ops, nnz_per_rank=None),
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": 1, "kind": kind,
                         "sample": f"columns [0,{block[4]}) of the workload ({nnz} stored entries), "
                                   f"{len(ops)} ops per step, serial (the reference hot path has no parallel pragma)"},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(spec, args, ops, nnz_per_rank):
    return {"workload": spec.name, "nrow": spec.nrow, "ncol_per_gpu": spec.ncol, "nnz_per_gpu": nnz_per_rank,
            "ops_per_step": list(ops), "values": "f64", "indices": "int32", "seed": spec.seed,
            "parallelism": f"column-sharded x{args.gpus}" if args.gpus > 1 else "single GPU",
            "l2": "inputs (1.2 GB/rank at C2) exceed the 126 MB L2; no explicit flush"}


# =====================================================================================================
# this repo's arm
# =====================================================================================================
def run_b200(args, ops):
    import torch
    import torch.distributed as dist

    from rcppsparse_b200 import DeviceMatrix, _lib, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: bench.py measures the CUDA path and has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    spec = synth.config(args.workload, args.scale)
    c0, c1 = rank * spec.ncol, (rank + 1) * spec.ncol  # weak scaling: every rank generates its own block
    D = DeviceMatrix.synth(spec, c0, c1, device=local_rank)
    nnz = D.nnz
    local = shard.GpuLocal(D)  # binds the handle to torch's current stream
    bounds = [k * spec.ncol for k in range(world + 1)]
    S = shard.ShardedMatrix(local, bounds, rank, device=dev, exchange=args.exchange)
    v_col = torch.empty(S.ncol, dtype=torch.float64, device=dev)
    v_row = torch.empty(S.nrow, dtype=torch.float64, device=dev)
    D.synth_vector_dev(spec.seed, 0, S.ncol, v_col)
    D.synth_vector_dev(spec.seed + 7, 0, S.nrow, v_row)
    T_keep = []

    # p2p: the exchange kernels of op k run on the window's stream beside the sweep of op k+1 (always);
    # nccl: only with --overlap (measured slower: the NCCL kernel waits for SM resources behind the sweep)
    overlap = world > 1 and (S.exchange == "p2p" or args.overlap)

    def run_op(op):
        """Launch op; with more than one rank its collective is left in flight (async NCCL) so that the next
        op's sweep overlaps it — every step waits for all of its results before it ends."""
        if op == "spmv":
            return S.spmv(v_col, async_op=overlap)
        if op == "spmv_t":
            return S.spmv_t(v_row, async_op=overlap)
        if op == "transpose":
            T_keep.clear()
            T_keep.append(D.transpose_dev())  # local block only (the sharded exchange is ShardedMatrix.transpose)
            return None
        return getattr(S, op)(async_op=overlap)

    def run_step(marks=None):
        pend = []
        for k, op in enumerate(ops):
            pend.append(run_op(op))
            if marks is not None:
                marks[k + 1].record()
        if overlap:
            for r in pend:
                if r is not None:
                    r.wait()
        if marks is not None:
            marks[len(ops) + 1].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- row sums of a resident mirror -------------------------------------------------------------------------
    # The library serves rowSums/rowMeans with its scatter kernel until a mirror has been asked for them more than
    # SB200_ROW_COMPANION_AFTER (8) times, then from a row-ordered copy of x (sparse_b200.h).  The steady state of
    # this benchmark is the second regime; the first is timed here, with the one-off cost of the switch, so that
    # the line carries both.
    row_companion = None
    if any(op in ("rowSums", "rowMeans") for op in ops):
        out_r = torch.empty(max(D.nrow, 1), dtype=torch.float64, device=dev)
        if args.no_row_companion:
            D.row_companion(-1)
        path0 = D.row_path()
        for _ in range(2):
            D.row_sums_dev(out_r)
        torch.cuda.synchronize()
        pre = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            D.row_sums_dev(out_r)
            e1.record()
            e1.synchronize()
            pre.append(e0.elapsed_time(e1))
        ab0 = D.algorithmic_bytes("row_sums")
        row_companion = {"scatter_path": path0, "scatter_ms_per_call": float(np.median(pre)),
                         "scatter_GBps": ab0 / (float(np.median(pre)) * 1e-3) / 1e9, "scatter_algorithmic_bytes": ab0,
                         "policy": "row-ordered copy after 8 row-sum calls on a mirror that owns its arrays",
                         "built": False}
        if not args.no_row_companion:
            try:
                t0 = time.perf_counter()
                D.row_companion(1)  # what the 9th call would do on its own; explicit so that it is timed, and so that
                torch.cuda.synchronize()  # it precedes the warm-up whatever --warmup is
                row_companion.update(built=True, build_ms=(time.perf_counter() - t0) * 1e3, extra_hbm_bytes=12 * nnz + 4 * (D.nrow + 1))
            except Exception as e:
                row_companion["error"] = f"{type(e).__name__}: {e}"
        del out_r

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        run_step()
    barrier()
    if rank == 0:
        sampler.wait_first_sample()
    launches0 = _lib.lib().sb200_launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 2)] for _ in range(args.steps)]
    barrier()
    sampler.mark_begin()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        ev[s][0].record()
        run_step(ev[s])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.mark_end()
    launches = _lib.lib().sb200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    total_ms = ev[0][0].elapsed_time(ev[-1][-1])  # device time of exactly K steps on the launching stream
    per_op_ms = {op: float(np.mean([ev[s][k].elapsed_time(ev[s][k + 1]) for s in range(args.steps)]))
                 for k, op in enumerate(ops)}
    per_op_min = {op: float(np.min([ev[s][k].elapsed_time(ev[s][k + 1]) for s in range(args.steps)]))
                  for k, op in enumerate(ops)}
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    nnz_all = torch.tensor([nnz], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(nnz_all, op=dist.ReduceOp.SUM)
    nnz_total = int(nnz_all.item())

    if world > 1:
        S.close()  # raises if an exchange barrier ever timed out
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    ms_per_step = total_ms / args.steps
    value = len(ops) * nnz_total / (ms_per_step * 1e-3)
    per_op = {}
    on_copy = D.row_path() == "row-companion"
    for op in ops:
        ab = D.algorithmic_bytes(ABI_OP[op] + ("_companion" if on_copy and op in ("rowSums", "rowMeans") else ""))
        gbs = ab / (per_op_ms[op] * 1e-3) / 1e9
        per_op[op] = {"ms": per_op_ms[op], "ms_min": per_op_min[op], "nnz_per_s_per_gpu": nnz / (per_op_ms[op] * 1e-3),
                      "algorithmic_bytes": ab, "achieved_GBps": gbs, "frac_of_measured": gbs / peak,
                      "frac_of_nominal_8TBps": gbs / NOMINAL_HBM_GBS}
    dom = max(ops, key=lambda o: per_op_ms[o])
    row_kernel = ("sweep_kernel<COLSUM> over the row-ordered copy" if on_copy else
                  "band_scatter_kernel" if D.row_path() == "banded" else "rowsum_stream_kernel")
    dom_kernel = {"rowSums": row_kernel, "rowMeans": row_kernel, "colSums": "sweep_kernel<COLSUM>",
                  "colMeans": "sweep_kernel<COLSUM>", "spmv": "sweep_kernel<SPMV>", "spmv_t": "sweep_kernel<SPMV_T>",
                  "transpose": "transpose_band_kernel"}[dom]
    roofline = {"bound": "hbm", "kernel": dom_kernel, "op": dom, "achieved": per_op[dom]["achieved_GBps"], "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": per_op[dom]["achieved_GBps"] / peak,
                "frac_of_nominal_8TBps": per_op[dom]["achieved_GBps"] / NOMINAL_HBM_GBS,
                "algorithmic_bytes_per_launch": per_op[dom]["algorithmic_bytes"], "ms_per_launch": per_op_ms[dom],
                "timed": "CUDA events on the launching stream around the op (zero-fill + kernel), mean over the timed steps",
                "traffic": profile_traffic(dom_kernel) if (args.workload, args.scale) == ("C2", 1.0) else None}
    if row_companion is not None:
        row_companion["scatter_frac_of_measured"] = row_companion["scatter_GBps"] / peak
        if world == 1 and row_companion.get("built"):
            row_companion["value_before_row_copy"] = value_before_row_copy(ops, nnz, per_op_ms, row_companion["scatter_ms_per_call"])

    line = {
        "metric": METRIC, "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(spec, args, ops, nnz), "per_op": per_op, "roofline": roofline,
        "row_path": D.row_path(), "row_companion": row_companion,
        "exchange": (None if world == 1 else
                     "libsparse_b200 kernels over NVLink peer memory (cudaIpc window): P2P stores of each rank's slice for "
                     "column results, rank-ordered P2P reduction for row results, flag barriers; op k's exchange runs beside op k+1's "
                     "sweep, every result is waited for before its step ends" if S.exchange == "p2p" else
                     "NCCL all-gather / all-reduce" + (", left in flight under the next op's sweep" if overlap else "")),
        "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
    }

    # ---- e2e: the host-buffer C ABI, upload + results every step (N = 1 only) -------------------------------
    if world == 1 and not args.no_e2e:
        try:
            line["e2e"] = measure_e2e(args, ops, D, nnz, local_rank)
        except Exception as e:  # never lose the device-resident line
            line["e2e"] = {"value": None, "unit": "nnz/s", "error": f"{type(e).__name__}: {e}"}
    elif world > 1:
        line["e2e"] = {"value": None, "unit": "nnz/s", "note": "measured at N=1 (host-buffer C ABI is per process)"}

    # ---- cpu_baseline: the reference's serial code on a bounded block, rank 0, N = 1 only ------------------------
    if world == 1 and not args.no_cpu_baseline:
        try:
            block = cpu_block(spec, args.cpu_cols)
            kind, t_cpu, nnz_cpu = time_cpu(ops, block, 3)
            cpu_step = sum(t_cpu.values())
            line["cpu_baseline"] = {
                "value": len(ops) * nnz_cpu / cpu_step, "unit": "nnz/s", "cores": 1, "kind": kind,
                "host_cores_available": os.cpu_count(),
                "sample": f"columns [0,{block[4]}) of the workload ({nnz_cpu} stored entries), best of 3 per op, serial",
                "per_op_nnz_per_s": {op: nnz_cpu / t_cpu[op] for op in ops}}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "nnz/s", "error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_e2e(args, ops, D, nnz, device):
    """Same step through the reference-facing host-buffer entry points: every step uploads i/p/x from
    pinned host memory into a fresh mirror (what one .Call from R pays), runs the ops, and reads every
    result vector back to host."""
    import torch

    from rcppsparse_b200 import DeviceMatrix

    i, p, x = D.download_columns()
    hi = torch.from_numpy(i).pin_memory()
    hp = torch.from_numpy(p).pin_memory()
    hx = torch.from_numpy(x).pin_memory()
    del i, p, x
    v_c = np.ones(D.ncol)
    v_r = np.ones(D.nrow)
    host_fn = {"colSums": "col_sums", "rowSums": "row_sums", "colMeans": "col_means", "rowMeans": "row_means"}

    # results land in caller-owned pinned buffers, as the inputs leave from pinned buffers (the C ABI takes the
    # caller's pointer either way; fresh pageable vectors would add their first-touch page faults to every step)
    res = {op: torch.empty(D.ncol if op in ("colSums", "colMeans", "spmv_t") else D.nrow, dtype=torch.float64).pin_memory()
           for op in ops if op != "transpose"}

    def step():
        with DeviceMatrix.from_host(hi, hp, hx, D.nrow, D.ncol, device=device, validate=True) as M:
            outs = []
            for op in ops:
                if op in host_fn:
                    outs.append(getattr(M, host_fn[op])(out=res[op]))
                elif op == "spmv":
                    outs.append(M.spmv(v_c))
                elif op == "spmv_t":
                    outs.append(M.spmv_t(v_r))
                elif op == "transpose":
                    outs.append(M.transpose_host()[1])
            return outs

    step()
    torch.cuda.synchronize()
    times = []
    for _ in range(args.e2e_steps):
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    h2d = 12 * nnz + 4 * (D.ncol + 1)
    d2h = sum(8 * (D.ncol if op in ("colSums", "colMeans", "spmv_t") else D.nrow) for op in ops if op != "transpose")
    out = {"value": len(ops) * nnz / dt, "unit": "nnz/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps, "ms_all_steps": [round(t * 1e3, 3) for t in times],
           "what": "per step: sb200_matrix_create from pinned host buffers (values upload overlapped with structure check "
                   "and plans) + the host-buffer ops (results copied back into pinned host buffers) + destroy; wall clock, "
                   "median step"}
    # the same step from PAGEABLE arrays (what an R caller owns): staged through pinned chunks by worker threads
    try:
        pi, pp, px = hi.numpy().copy(), hp.numpy().copy(), hx.numpy().copy()
        pres = {op: np.empty(res[op].shape[0]) for op in res}

        def pstep():
            with DeviceMatrix.from_host(pi, pp, px, D.nrow, D.ncol, device=device, validate=True) as M:
                for op in ops:
                    if op in host_fn:
                        getattr(M, host_fn[op])(out=pres[op])

        pstep()
        ptimes = []
        for _ in range(max(2, args.e2e_steps // 2)):
            t0 = time.perf_counter()
            pstep()
            torch.cuda.synchronize()
            ptimes.append(time.perf_counter() - t0)
        pdt = float(np.median(ptimes))
        out["pageable"] = {"value": len(ops) * nnz / pdt, "ms_per_step": pdt * 1e3,
                           "what": "same step with pageable host arrays for inputs and results"}
    except Exception as e:
        out["pageable"] = {"value": None, "error": f"{type(e).__name__}: {e}"}
    return out


def main():
    # NCCL prints a version banner on stdout at init when NCCL_DEBUG=VERSION/INFO is in the environment;
    # rank 0 must print exactly one JSON line
    os.environ["NCCL_DEBUG"] = os.environ.get("SB200_NCCL_DEBUG", "WARN")
    args = parse_args()
    ops = tuple(o for o in args.ops.split(",") if o)
    for o in ops:
        if o not in ABI_OP:
            raise SystemExit(f"unknown op {o}")
    if args.impl == "reference":
        run_reference(args, tuple(o for o in ops if o != "transpose") or OPS)
    else:
        run_b200(args, ops)


if __name__ == "__main__":
    main()
