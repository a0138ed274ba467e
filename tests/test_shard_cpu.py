"""Host-side logic of the column-sharding layer at world_size 2 over gloo (no GPU):
the split, the slice bookkeeping and the collectives.  The rank-local compute is an
oracle-backed stand-in for the CUDA kernels, injected through the LocalSweeps protocol."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from rcppsparse_b200 import shard, synth


class OracleLocal:
    """LocalSweeps on CPU tensors, computed by the C port (test stand-in for GpuLocal)."""

    def __init__(self, i, p, x, nrow, ncol):
        self.i, self.p, self.x, self.nrow, self.ncol = i, p, x, nrow, ncol
        self.o = oracle.Port()

    def col_sums(self, out, divisor):
        r = self.o.colSums(self.i, self.p, self.x, self.nrow, self.ncol)
        out.copy_(torch.from_numpy(r / divisor if divisor else r))

    def row_sums(self, out):
        out.copy_(torch.from_numpy(self.o.rowSums(self.i, self.p, self.x, self.nrow, self.ncol)))

    def spmv(self, v_local, out):
        out.copy_(torch.from_numpy(self.o.spmv(self.i, self.p, self.x, self.nrow, self.ncol, v_local.numpy())))

    def spmv_t(self, v, out):
        out.copy_(torch.from_numpy(self.o.spmv_t(self.i, self.p, self.x, self.nrow, self.ncol, v.numpy())))

    def div(self, t, divisor):
        t.div_(divisor)

    def transpose_local(self):
        ti, tp, tx = self.o.transpose(self.i, self.p, self.x, self.nrow, self.ncol)
        return torch.from_numpy(tp), torch.from_numpy(ti), torch.from_numpy(tx)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, balanced, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = synth.powerlaw_spec(700, 240, 30.0, 99, row_levels=4)
        i, p, x = synth.generate_host(spec)
        bounds = shard.split_columns_by_nnz(p, world) if balanced else shard.split_columns_evenly(spec.ncol, world)
        c0, c1 = bounds[rank], bounds[rank + 1]
        # a rank's block can be generated independently of the others (weak-scaling benches do this)
        bi, bp, bx = synth.generate_host(spec, c0, c1)
        assert np.array_equal(bp, p[c0:c1 + 1] - p[c0]) and np.array_equal(bi, i[p[c0]:p[c1]])
        S = shard.ShardedMatrix(OracleLocal(bi, bp, bx, spec.nrow, c1 - c0), bounds, rank)
        full = oracle.Port()
        args = (i, p, x, spec.nrow, spec.ncol)
        v_c = torch.from_numpy(synth.dense_vector(1, spec.ncol))
        v_r = torch.from_numpy(synth.dense_vector(2, spec.nrow))
        checks = {
            "colSums": (S.colSums().numpy().copy(), full.colSums(*args), None),
            "colMeans": (S.colMeans().numpy().copy(), full.colMeans(*args), None),
            "rowSums": (S.rowSums().numpy().copy(), full.rowSums(*args), None),
            "rowMeans": (S.rowMeans().numpy().copy(), full.rowMeans(*args), None),
            "spmv": (S.spmv(v_c).numpy().copy(), full.spmv(*args, v_c.numpy()), v_c.numpy()),
            "spmv_t": (S.spmv_t(v_r).numpy().copy(), full.spmv_t(*args, v_r.numpy()), v_r.numpy()),
        }
        for op, (got, want, v) in checks.items():
            oracle.assert_within(op, got, want, *args, v=v)
        # the same six sweeps launched back to back with their collectives left in flight (what bench.py does
        # so that op k's collective overlaps op k+1's sweep), waited afterwards in a different order
        pend = {"colSums": S.colSums(async_op=True), "rowSums": S.rowSums(async_op=True),
                "colMeans": S.colMeans(async_op=True), "rowMeans": S.rowMeans(async_op=True),
                "spmv": S.spmv(v_c, async_op=True), "spmv_t": S.spmv_t(v_r, async_op=True)}
        for op in ("rowMeans", "spmv_t", "colSums", "spmv", "colMeans", "rowSums"):
            got = pend[op].wait().numpy()
            oracle.assert_within(op, got, checks[op][1], *args, v=checks[op][2])
            assert pend[op].wait() is pend[op].tensor  # a second wait is a no-op (no second division)
            oracle.assert_within(op, pend[op].tensor.numpy(), checks[op][1], *args, v=checks[op][2])
        # sharded transpose: my row block of A^T must equal the same rows of the full transpose, bit for bit
        rb, tp_own, tcols, tvals = S.transpose()
        fi, fp, fx = full.transpose(*args)
        r0, r1 = rb[rank], rb[rank + 1]
        assert rb[0] == 0 and rb[-1] == spec.nrow and all(rb[k] <= rb[k + 1] for k in range(world))
        assert np.array_equal(tp_own.numpy(), fp[r0:r1 + 1] - fp[r0])
        assert np.array_equal(tcols.numpy(), fi[fp[r0]:fp[r1]])
        assert np.array_equal(tvals.numpy().view(np.uint64), fx[fp[r0]:fp[r1]].view(np.uint64))
        q.put((rank, "ok", bounds))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, f"FAIL {type(e).__name__}: {e}", None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balanced", [True, False])
def test_sharded_sweeps_world2_gloo(balanced):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, balanced, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results


def test_split_by_nnz_balances_and_covers():
    spec = synth.powerlaw_spec(5000, 2000, 80.0, 5)
    _, p, _ = synth.generate_host(spec)
    for world in (1, 2, 3, 8):
        b = shard.split_columns_by_nnz(p, world)
        assert b[0] == 0 and b[-1] == spec.ncol and all(b[k] <= b[k + 1] for k in range(world))
        per = [int(p[b[k + 1]] - p[b[k]]) for k in range(world)]
        assert sum(per) == int(p[-1])
        longest = int(np.diff(p).max())
        assert max(per) - min(per) <= 2 * longest + 1  # never splits a column, so off by at most one column each side


def test_split_handles_degenerate_matrices():
    assert shard.split_columns_by_nnz(np.zeros(1, np.int32), 4) == [0, 0, 0, 0, 0]       # ncol = 0
    assert shard.split_columns_by_nnz(np.zeros(6, np.int32), 2) == [0, 0, 5]             # all columns empty
    assert shard.split_columns_evenly(10, 4) == [0, 2, 5, 7, 10]


class _FakeExchangeLib:
    """Stands in for libsparse_b200's exchange entry points so that PeerWindow's handshake can run without a GPU."""

    def __init__(self, create_rc=0, connect_rc=0):
        self.create_rc, self.connect_rc, self.calls = create_rc, connect_rc, []

    def sb200_exchange_create(self, device, nbytes, out_handle, ipc):
        self.calls.append("create")
        if self.create_rc == 0:
            out_handle._obj.value = 0x1234
        return self.create_rc

    def sb200_exchange_connect(self, h, rank, world, blob):
        self.calls.append(("connect", rank, world, len(blob)))
        return self.connect_rc

    def sb200_exchange_window(self, h, base, off, nbytes):
        base._obj.value, off._obj.value, nbytes._obj.value = 1 << 20, 4096, 1 << 16
        return 0

    def sb200_exchange_destroy(self, h):
        self.calls.append("destroy")
        return 0

    def sb200_last_error(self):
        return b"fake failure"


@pytest.mark.parametrize("create_rc,connect_rc,ok", [(0, 0, True), (5, 0, False), (0, 5, False)])
def test_peer_window_handshake_never_strands_a_rank(create_rc, connect_rc, ok, monkeypatch):
    """Whatever fails locally (creating the window, mapping a peer), a rank still takes part in both collectives of
    the handshake and then every rank raises together — nobody is left waiting in a collective."""
    from rcppsparse_b200 import _lib

    fake = _FakeExchangeLib(create_rc, connect_rc)
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        if ok:
            W = shard.PeerWindow(torch.device("cpu"), 0, 1, 1 << 12)
            assert fake.calls[:2] == ["create", ("connect", 0, 1, 64)] and W.bytes == 1 << 16
        else:
            with pytest.raises(RuntimeError, match="peer window not available"):
                shard.PeerWindow(torch.device("cpu"), 0, 1, 1 << 12)
            assert "destroy" in fake.calls or create_rc != 0
    finally:
        dist.destroy_process_group()
