#!/usr/bin/env python
"""Cost of the exchange kernels alone (torchrun, one rank per GPU): barrier, gather of a column result,
reduction of a row result — back to back, CUDA events on rank 0, against the NCCL collectives of the same size.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/xg_bench.py [--ncol 100000 --nrow 1000000]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rcppsparse_b200 import shard  # noqa: E402


def timed(fn, reps, dev):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncol", type=int, default=100000, help="columns per rank")
    ap.add_argument("--nrow", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=200)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ncol = a.ncol * world
    W = shard.PeerWindow(dev, rank, world, 8 * (ncol + 2 * a.nrow) + 4096)
    coff, cfull = W.alloc(ncol)
    poff, part = W.alloc(a.nrow)
    roff, res = W.alloc(a.nrow)
    cfull.fill_(float(rank))
    part.fill_(1.0 + rank)
    torch.cuda.synchronize()
    dist.barrier()
    main = torch.cuda.current_stream()
    out = {"barrier_us": timed(lambda: main.wait_event(W.barrier()), a.reps, dev),
           "gather_us": timed(lambda: main.wait_event(W.gather(coff, rank * a.ncol, a.ncol)), a.reps, dev),
           "reduce_us": timed(lambda: main.wait_event(W.reduce(poff, roff, a.nrow, 0.0)), a.reps, dev)}
    W.status()
    assert float(res[0]) == sum(1.0 + q for q in range(world)) and float(res[-1]) == float(res[0])
    assert all(float(cfull[q * a.ncol]) == float(q) for q in range(world))
    g_in = torch.ones(a.ncol, dtype=torch.float64, device=dev)
    g_out = torch.empty(ncol, dtype=torch.float64, device=dev)
    r = torch.ones(a.nrow, dtype=torch.float64, device=dev)
    out["nccl_all_gather_us"] = timed(lambda: dist.all_gather_into_tensor(g_out, g_in), a.reps, dev)
    out["nccl_all_reduce_us"] = timed(lambda: dist.all_reduce(r), a.reps, dev)
    if rank == 0:
        print({"world": world, "ncol_per_rank": a.ncol, "nrow": a.nrow, **{k: round(v, 1) for k, v in out.items()}})
    dist.barrier()
    W.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
