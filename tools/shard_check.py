#!/usr/bin/env python
"""Multi-GPU parity check (torchrun, NCCL): every op of the column-sharded matrix against the oracle on
the full matrix, on every rank.  Small matrix so the CPU oracle takes a moment."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from rcppsparse_b200 import DeviceMatrix, shard, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    worst = {}
    used = set()
    for name, spec in (("C2/50", synth.config("C2", 0.02)), ("C3/250", synth.config("C3", 0.004)),
                       ("powerlaw", synth.powerlaw_spec(40_000, 9_001, 150.0, 31))):
        i, p, x = synth.generate_host(spec)
        for balanced in (True, False):
            bounds = shard.split_columns_by_nnz(p, world) if balanced else shard.split_columns_evenly(spec.ncol, world)
            c0, c1 = bounds[rank], bounds[rank + 1]
            D = DeviceMatrix.synth(spec, c0, c1, device=local)  # each rank generates only its block
            # the library's peer-memory exchange on the balanced split, NCCL collectives on the even one
            S = shard.ShardedMatrix(shard.GpuLocal(D), bounds, rank, device=dev, exchange="p2p" if balanced else "nccl")
            used.add(S.exchange)
            chk = oracle.best()
            args = (i, p, x, spec.nrow, spec.ncol)
            v_c = torch.from_numpy(synth.dense_vector(1, spec.ncol)).to(dev)
            v_r = torch.from_numpy(synth.dense_vector(2, spec.nrow)).to(dev)
            # every method returns the object's internal result buffer: copy before the next call
            for op, run, want, v in (
                    ("colSums", S.colSums, chk.colSums(*args), None), ("colMeans", S.colMeans, chk.colMeans(*args), None),
                    ("rowSums", S.rowSums, chk.rowSums(*args), None), ("rowMeans", S.rowMeans, chk.rowMeans(*args), None),
                    ("spmv", lambda: S.spmv(v_c), chk.spmv(*args, v_c.cpu().numpy()), v_c.cpu().numpy()),
                    ("spmv_t", lambda: S.spmv_t(v_r), chk.spmv_t(*args, v_r.cpu().numpy()), v_r.cpu().numpy())):
                for rep in range(3):  # repeated calls walk the alternating result buffers of the window
                    got = run().cpu().numpy()
                    if rep == 0:
                        first = got
                    elif S.exchange == "p2p":
                        assert np.array_equal(got.view(np.uint64), first.view(np.uint64)) or op in ("rowSums", "rowMeans", "spmv"), \
                            f"{op}: column results must be bit-stable call to call"
                    r = oracle.assert_within(op, got, want, *args, v=v)
                    worst[op] = max(worst.get(op, 0.0), r)
                # every rank holds the same bits (the row reduction adds in rank order on the owner of each row block)
                if world > 1:
                    mine = run().clone()
                    ref = mine.clone()
                    dist.broadcast(ref, src=0)
                    assert torch.equal(mine.view(torch.int64), ref.view(torch.int64)), f"{op}: ranks disagree bitwise"
            # sharded transpose: local device transposes + one exchange step (P2P segment pushes on the peer-memory
            # exchange, an NCCL all-to-all-v on the collective one): my rows, bit for bit
            rb, tp_own, tcols, tvals = S.transpose()
            fi, fp, fx = chk.transpose(*args)
            r0, r1 = rb[rank], rb[rank + 1]
            assert np.array_equal(tp_own.cpu().numpy(), fp[r0:r1 + 1] - fp[r0]), "sharded transpose: p differs"
            assert np.array_equal(tcols.cpu().numpy(), fi[fp[r0]:fp[r1]]), "sharded transpose: column ids differ"
            assert np.array_equal(tvals.cpu().numpy().view(np.uint64), fx[fp[r0]:fp[r1]].view(np.uint64))
            S.close()
            D.close()
    dist.barrier()
    print(f"rank {rank}/{world}: sharded parity ok, exchange {sorted(used)} (sums, SpMV, transpose bit-exact), worst |err|/sum|a| " + ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
