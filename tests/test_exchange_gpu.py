"""The cross-GPU exchange layer (sparse_b200.h, sb200_exchange_*; rcppsparse_b200/shard.py PeerWindow).

World size 1 runs on any GPU box (window, offsets, the reduction kernel's own arithmetic); the multi-rank
check needs two GPUs and runs tools/shard_check.py under torchrun (every sharded op against the oracle over
the peer-memory exchange and over NCCL; `gpurun --gpus 2 -- python -m pytest tests/test_exchange_gpu.py -m gpu`).
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_window_world1_reduce_gather_and_bounds():
    import torch
    import torch.distributed as dist

    from rcppsparse_b200 import SparseB200Error, shard

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1, device_id=dev)
    try:
        n = 100_003  # odd: the scalar tail row of the reduction kernel
        W = shard.PeerWindow(dev, 0, 1, 8 * 3 * n + 4096)
        poff, part = W.alloc(n)
        roff, res = W.alloc(n)
        coff, col = W.alloc(n)
        assert poff % 256 == 0 and roff % 256 == 0 and roff >= poff + 8 * n
        part.copy_(torch.arange(n, dtype=torch.float64, device=dev) * 0.25 + 1.0)
        res.fill_(-1.0)
        main = torch.cuda.current_stream()
        main.wait_event(W.reduce(poff, roff, n, 0.0))
        assert torch.equal(res, part)
        main.wait_event(W.reduce(poff, roff, n, 7.0))
        # IEEE division, as the mean's sums / Dim (numpy divides; torch multiplies by the scalar's reciprocal on CUDA)
        assert np.array_equal(res.cpu().numpy(), part.cpu().numpy() / 7.0)
        col.fill_(3.0)
        main.wait_event(W.gather(coff, 10, n - 10))  # world 1: nothing to push, nothing to wait for
        main.wait_event(W.barrier())
        W.status()
        assert float(col.sum()) == 3.0 * n
        with pytest.raises(SparseB200Error):
            W.reduce(poff, roff + 8 * n, n, 0.0)  # result would run past the window
        with pytest.raises(MemoryError):
            W.alloc(n)
        W.close()
    finally:
        dist.destroy_process_group()


def test_sharded_matrix_world1_uses_plain_buffers():
    """With one rank there is no exchange: ShardedMatrix must not need a window (or a process group)."""
    import torch

    from oracle import oracle
    from rcppsparse_b200 import DeviceMatrix, shard, synth

    spec = synth.config("C1")
    i, p, x = synth.generate_host(spec)
    args = (i, p, x, spec.nrow, spec.ncol)
    with DeviceMatrix.synth(spec) as D:
        S = shard.ShardedMatrix(shard.GpuLocal(D), [0, spec.ncol], 0, device=torch.device("cuda", 0))
        assert S.exchange == "collective" and S.window is None
        chk = oracle.best()
        oracle.assert_within("rowMeans", S.rowMeans().cpu().numpy(), chk.rowMeans(*args), *args, tol=1e-12)
        oracle.assert_within("colSums", S.colSums(async_op=True).wait().cpu().numpy(), chk.colSums(*args), *args, tol=1e-12)
        S.close()


def test_two_ranks_peer_memory_exchange_against_oracle():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one node (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "shard_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("sharded parity ok") == 2 and "'p2p'" in r.stdout
