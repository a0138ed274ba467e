"""Column sharding of one dgCMatrix across the ranks of a torch.distributed group.

The reference is single-process (no parallelism on the hot path); the
sweeps shard naturally because columns are independent units:

  * each rank holds a contiguous, nnz-balanced COLUMN BLOCK — itself a valid dgCMatrix with
    the full row count and p rebased to 0 — as one device-resident mirror (DeviceMatrix);
  * column-indexed results (colSums, colMeans, A^T v) are disjoint slices: no arithmetic across
    ranks, one all-gather to assemble the vector on every rank;
  * row-indexed results (rowSums, rowMeans, A v) are full-length partials: one all-reduce (sum)
    of 8*nrow bytes; mean scaling by the GLOBAL ncol happens after the reduce.

One process per GPU (torchrun); NCCL over NVLink on GPUs, gloo in the CPU tests.  The local
compute is behind a tiny protocol (``LocalSweeps``) so the host-side logic — the split, the
slice bookkeeping, the collectives — is testable at world_size 2 without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Protocol, Sequence

import numpy as np
import torch
import torch.distributed as dist


def split_columns_by_nnz(p: np.ndarray, world: int) -> list[int]:
    """Column boundaries b[0..world] with p[b[k]] ~= k*nnz/world (binary search in p).
    A column is never split; boundaries are non-decreasing; b[0]=0, b[world]=ncol."""
    ncol = int(p.shape[0]) - 1
    nnz = int(p[ncol])
    bounds = [0]
    for k in range(1, world):
        target = (nnz * k) // world
        c = int(np.searchsorted(p, target, side="left"))
        c = min(max(c, bounds[-1]), ncol)
        bounds.append(c)
    bounds.append(ncol)
    return bounds


def split_columns_evenly(ncol: int, world: int) -> list[int]:
    return [(ncol * k) // world for k in range(world + 1)]


class LocalSweeps(Protocol):
    """What a rank's column block must offer; tensors live on the rank's device."""

    nrow: int
    ncol: int

    def col_sums(self, out: torch.Tensor, divisor: float) -> None: ...
    def row_sums(self, out: torch.Tensor) -> None: ...
    def spmv(self, v_local: torch.Tensor, out: torch.Tensor) -> None: ...
    def spmv_t(self, v: torch.Tensor, out: torch.Tensor) -> None: ...
    def div(self, t: torch.Tensor, divisor: float) -> None: ...
    def transpose_local(self): ...  # -> (p int32[nrow+1], cols int32[nnz] local ids, vals float64[nnz]) on the device


class _CudaView:
    """__cuda_array_interface__ shim: a raw device pointer as a tensor, no copy."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class GpuLocal:
    """LocalSweeps over a DeviceMatrix: every method is a kernel launch on torch's current stream."""

    def __init__(self, dm):
        self.dm = dm
        self.nrow, self.ncol = dm.nrow, dm.ncol
        dm.set_stream(torch.cuda.current_stream().cuda_stream)

    def col_sums(self, out, divisor):
        self.dm.col_sums_dev(out, divisor)

    def row_sums(self, out):
        self.dm.row_sums_dev(out, 0.0)

    def spmv(self, v_local, out):
        self.dm.spmv_dev(v_local, out)

    def spmv_t(self, v, out):
        self.dm.spmv_t_dev(v, out)

    def div(self, t, divisor):
        self.dm.vec_div_dev(t, t.numel(), divisor)

    def transpose_local(self):
        """Transpose this rank's column block on the device; the result arrays are viewed as torch tensors
        without copying (the DeviceMatrix that owns them is kept alive alongside)."""
        T = self.dm.transpose_dev()
        T.sync()
        pi, pp, px = T.device_arrays()
        dev = torch.device("cuda", torch.cuda.current_device())
        p = torch.as_tensor(_CudaView(pp, T.ncol + 1, "<i4"), device=dev)
        if T.nnz > 0:
            i = torch.as_tensor(_CudaView(pi, T.nnz, "<i4"), device=dev)
            x = torch.as_tensor(_CudaView(px, T.nnz, "<f8"), device=dev)
        else:
            i = torch.empty(0, dtype=torch.int32, device=dev)
            x = torch.empty(0, dtype=torch.float64, device=dev)
        self._keep_t = T
        return p, i, x


class PeerWindow:
    """A rank's window of device memory mapped by every rank of the node (libsparse_b200's exchange layer,
    sparse_b200.h): result vectors of a ShardedMatrix live inside it and the two exchange steps of the path are
    the library's own kernels over NVLink peer memory — P2P stores of a rank's slice for column-indexed results,
    a rank-ordered P2P reduction of the partials for row-indexed ones — instead of NCCL calls.
    torch.distributed only carries the 64-byte handles (and agrees on whether every rank could connect)."""

    def __init__(self, device: torch.device, rank: int, world: int, data_bytes: int, group=None):
        from . import _lib
        self._lib, self.device, self.rank, self.world = _lib, device, rank, world
        self._h = C.c_void_p()
        self.side = None
        handle = (C.c_ubyte * 64)()
        # every rank goes through both collectives below whatever happened locally, so a rank that cannot create
        # or map a window never leaves the others waiting
        rc = _lib.lib().sb200_exchange_create(device.index, int(data_bytes) + 4096, C.byref(self._h), handle)
        err = _lib.lib().sb200_last_error().decode(errors="replace") if rc != 0 else ""
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(handle) if rc == 0 else None, group=group)
        if rc == 0 and all(g is not None for g in gathered):
            rc = _lib.lib().sb200_exchange_connect(self._h, rank, world, b"".join(gathered))
            err = _lib.lib().sb200_last_error().decode(errors="replace") if rc != 0 else ""
        elif rc == 0:
            rc, err = -1, "a peer could not create its window"
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:  # some rank could not map its peers: nobody uses the window
            self.close()
            raise RuntimeError(f"peer window not available on every rank (rank {rank}: {err or 'ok'})")
        base, off, nbytes = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().sb200_exchange_window(self._h, C.byref(base), C.byref(off), C.byref(nbytes)))
        self.base, self._cursor, self.bytes = base.value, off.value, nbytes.value

    def alloc(self, n: int):
        """(byte offset, float64 tensor of n entries) inside the window; 256-byte aligned."""
        off = self._cursor
        need = max(int(n), 1) * 8
        if off + need > self.bytes:
            raise MemoryError("peer window exhausted")
        self._cursor = (off + need + 255) & ~255
        t = torch.as_tensor(_CudaView(self.base + off, max(int(n), 1), "<f8"), device=self.device)[:n]
        return off, t

    # Exchange kernels run on the window's own high-priority stream, ordered after the caller's current stream
    # at the time of the call; each call returns the event that marks its completion.  They are a few small CTAs
    # without shared memory, so they run beside the next sweep (whose persistent CTAs fill the SMs' shared
    # memory but not their thread slots) instead of behind it; every rank issues them in the same order on
    # that one stream, which is what the epoch barriers need.
    def _run(self, launch) -> torch.cuda.Event:
        if self.side is None:
            self.side = torch.cuda.Stream(self.device, priority=-1)
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        self._lib.check(launch(C.c_void_p(self.side.cuda_stream)))
        ev = torch.cuda.Event()
        ev.record(self.side)
        return ev

    def gather(self, full_off: int, slice_begin: int, slice_len: int) -> torch.cuda.Event:
        return self._run(lambda st: self._lib.lib().sb200_exchange_gather(self._h, st, full_off, slice_begin, slice_len))

    def reduce(self, partial_off: int, result_off: int, n: int, divisor: float = 0.0) -> torch.cuda.Event:
        return self._run(lambda st: self._lib.lib().sb200_exchange_reduce(self._h, st, partial_off, result_off, n, divisor))

    def barrier(self) -> torch.cuda.Event:
        return self._run(lambda st: self._lib.lib().sb200_exchange_barrier(self._h, st))

    def status(self) -> None:
        """Synchronises; raises if a barrier gave up on a peer."""
        self._lib.check(self._lib.lib().sb200_exchange_status(self._h))

    def close(self) -> None:
        if self._h:
            self._lib.lib().sb200_exchange_destroy(self._h)
            self._h = C.c_void_p()


class Pending:
    """Result of a sweep whose collective may still be in flight (``async_op=True``): ``wait()`` orders the
    caller's current stream after it, applies what has to follow the collective (the mean's division) and
    returns the full vector.  The sweep of the next op can be launched before ``wait()``: its kernel and
    this op's collective then overlap, as in ``torch.distributed``'s own ``async_op``."""

    def __init__(self, tensor, works=(), finish=None):
        self.tensor = tensor
        self._works = [w for w in works if w is not None]
        self._finish = finish

    def wait(self) -> torch.Tensor:
        for w in self._works:
            w.wait()
        self._works = []
        if self._finish is not None:
            self._finish()
            self._finish = None
        return self.tensor


class ShardedMatrix:
    """The reference's Matrix methods (RcppSparse.h:131-156 + the SpMV idiom) over a column-sharded matrix.
    Every method returns the FULL result vector on every rank — as a view of an internal buffer that
    the next call of the SAME method overwrites: clone it to keep it.  With ``async_op=True`` a method
    returns a ``Pending`` instead; wait for it before calling the same method again."""

    def __init__(self, local: LocalSweeps, bounds: Sequence[int], rank: int, group=None, device=None,
                 exchange: str | None = None):
        """exchange: "p2p" (the library's kernels over a PeerWindow; CUDA ranks of one node), "nccl"/"gloo"
        (torch.distributed collectives), default: SB200_EXCHANGE or "p2p" on CUDA with more than one rank.
        Construction is collective when the window is used (handles are exchanged)."""
        self.local = local
        self.bounds = list(bounds)
        self.world = len(self.bounds) - 1
        self.rank = rank
        self.group = group
        self.nrow = local.nrow
        self.ncol = self.bounds[-1]
        self.c0, self.c1 = self.bounds[rank], self.bounds[rank + 1]
        if local.ncol != self.c1 - self.c0:
            raise ValueError("local block does not match its column range")
        self.device = device if device is not None else torch.device("cpu")
        self._counts = [self.bounds[k + 1] - self.bounds[k] for k in range(self.world)]
        self._even = len(set(self._counts)) == 1
        self._bufs: dict[str, torch.Tensor] = {}
        self.window = None
        self._xbuf: dict[str, list] = {}
        self._inflight: dict[str, Pending] = {}
        want = exchange or os.environ.get("SB200_EXCHANGE") or ("p2p" if self.device.type == "cuda" and self.world > 1 else "collective")
        if want == "p2p" and self.world > 1:
            if self.device.type != "cuda":
                raise ValueError("exchange='p2p' needs CUDA ranks")
            pad = 256
            data = 3 * 2 * (8 * self.ncol + pad) + 3 * 3 * (8 * self.nrow + pad)
            try:
                self.window = PeerWindow(self.device, rank, self.world, data, group)
            except RuntimeError as e:  # raised on EVERY rank or on none (the ranks agree inside PeerWindow)
                if exchange == "p2p" or os.environ.get("SB200_EXCHANGE") == "p2p":
                    raise
                import warnings
                warnings.warn(f"peer-memory exchange unavailable, using torch.distributed collectives: {e}")
        self.exchange = "p2p" if self.window is not None else "collective"

    def close(self) -> None:
        if self.window is not None:
            failure = None
            try:
                self.window.status()
            except Exception as e:  # reported after the collective below: the other ranks are waiting in it
                failure = e
            dist.barrier(group=self.group)  # nobody unmaps while a peer may still write
            self.window.close()
            self.window = None
            if failure is not None:
                raise failure

    # ---- p2p path: results live in the window ------------------------------------------------------------------
    def _window_cols(self, name: str):
        """(offset, tensor) of the result buffer to fill now: two per method, alternating — a peer that is one
        call ahead pushes its slice into the other one, never into the result the caller still holds."""
        st = self._xbuf.get(name)
        if st is None:
            st = self._xbuf[name] = [0, self.window.alloc(self.ncol), self.window.alloc(self.ncol)]
        st[0] ^= 1
        return st[1 + st[0]]

    def _window_rows(self, name: str):
        """((partial offset, tensor), (result offset, tensor)): one partial, two alternating results."""
        st = self._xbuf.get(name)
        if st is None:
            st = self._xbuf[name] = [0, self.window.alloc(self.nrow), self.window.alloc(self.nrow), self.window.alloc(self.nrow)]
        st[0] ^= 1
        return st[1], st[2 + st[0]]

    def _p2p_pending(self, name: str, tensor: torch.Tensor, ev: torch.cuda.Event) -> Pending:
        p = Pending(tensor, finish=lambda: torch.cuda.current_stream(self.device).wait_event(ev))
        self._inflight[name] = p
        return p

    def _p2p_cols(self, name: str, launch) -> Pending:
        prev = self._inflight.pop(name, None)
        if prev is not None:
            prev.wait()  # the previous call of this method must have landed before its buffers are reused
        off, full = self._window_cols(name)
        launch(full[self.c0:self.c1])
        return self._p2p_pending(name, full, self.window.gather(off, self.c0, self.c1 - self.c0))

    def _p2p_rows(self, name: str, launch, divisor: float = 0.0) -> Pending:
        prev = self._inflight.pop(name, None)
        if prev is not None:
            prev.wait()  # peers read my partial until that exchange has ended
        (poff, partial), (roff, result) = self._window_rows(name)
        launch(partial)
        return self._p2p_pending(name, result, self.window.reduce(poff, roff, self.nrow, divisor))

    def _buf(self, name: str, n: int) -> torch.Tensor:
        t = self._bufs.get(name)
        if t is None:
            t = self._bufs[name] = torch.empty(n, dtype=torch.float64, device=self.device)
        return t

    @staticmethod
    def _done(p: Pending, async_op: bool):
        return p if async_op else p.wait()

    # ---- column-indexed: disjoint slices + all-gather --------------------------------------------------
    def _gather_columns(self, full: torch.Tensor) -> Pending:
        if self.world == 1:
            return Pending(full)
        if self._even:
            mine = full[self.c0:self.c1].clone()  # kept alive by the Pending until the gather has run
            w = dist.all_gather_into_tensor(full, mine, group=self.group, async_op=True)
            p = Pending(full, [w])
            p._keep = mine
            return p
        works = []
        # uneven counts: one broadcast per owner (grouped by the backend)
        for k in range(self.world):
            part = full[self.bounds[k]:self.bounds[k + 1]]
            if part.numel():
                works.append(dist.broadcast(part, src=dist.get_global_rank(self.group, k) if self.group else k,
                                            group=self.group, async_op=True))
        return Pending(full, works)

    def colSums(self, async_op: bool = False):
        if self.window is not None:
            return self._done(self._p2p_cols("colSums", lambda out: self.local.col_sums(out, 0.0)), async_op)
        full = self._buf("colSums", self.ncol)
        self.local.col_sums(full[self.c0:self.c1], 0.0)
        return self._done(self._gather_columns(full), async_op)

    def colMeans(self, async_op: bool = False):
        if self.window is not None:
            return self._done(self._p2p_cols("colMeans", lambda out: self.local.col_sums(out, float(self.nrow))), async_op)
        full = self._buf("colMeans", self.ncol)
        self.local.col_sums(full[self.c0:self.c1], float(self.nrow))  # nrow is global already
        return self._done(self._gather_columns(full), async_op)

    def spmv_t(self, v: torch.Tensor, async_op: bool = False):
        """A^T v: v[nrow] replicated on every rank."""
        if self.window is not None:
            return self._done(self._p2p_cols("spmv_t", lambda out: self.local.spmv_t(v, out)), async_op)
        full = self._buf("spmv_t", self.ncol)
        self.local.spmv_t(v, full[self.c0:self.c1])
        return self._done(self._gather_columns(full), async_op)

    # ---- row-indexed: full-length partials + all-reduce ----------------------------------------------------
    def _reduce_rows(self, full: torch.Tensor, finish=None) -> Pending:
        works = []
        if self.world > 1:
            works.append(dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        return Pending(full, works, finish)

    def rowSums(self, async_op: bool = False):
        if self.window is not None:
            return self._done(self._p2p_rows("rowSums", self.local.row_sums), async_op)
        full = self._buf("rowSums", self.nrow)
        self.local.row_sums(full)
        return self._done(self._reduce_rows(full), async_op)

    def rowMeans(self, async_op: bool = False):
        if self.window is not None:  # the division by the GLOBAL ncol rides in the reduction kernel
            return self._done(self._p2p_rows("rowMeans", self.local.row_sums, float(self.ncol)), async_op)
        full = self._buf("rowMeans", self.nrow)
        self.local.row_sums(full)
        # GLOBAL ncol, after the reduce (RcppSparse.h:154 divides the finished sums by Dim[1])
        return self._done(self._reduce_rows(full, lambda: self.local.div(full, float(self.ncol))), async_op)

    def spmv(self, v: torch.Tensor, async_op: bool = False):
        """A v: v[ncol] replicated; each rank uses only its own slice."""
        if self.window is not None:
            return self._done(self._p2p_rows("spmv", lambda out: self.local.spmv(v[self.c0:self.c1], out)), async_op)
        full = self._buf("spmv", self.nrow)
        self.local.spmv(v[self.c0:self.c1], full)
        return self._done(self._reduce_rows(full), async_op)

    # ---- transpose: local transposes + ONE exchange step (all-to-all-v) --------------------------------------
    def transpose(self):
        """CSC of A^T, row-sharded.  Every rank transposes its own column block with the
        device kernel; row r of the result is the concatenation, in rank (= column) order, of each rank's
        row-r segment, so order inside a row is ascending source column by construction.  Rows are then
        dealt out in nnz-balanced contiguous blocks: one all-gather of the per-rank row counts, one
        all-to-all-v of (column id, value) pairs, and an index gather that interleaves the received
        per-rank segments row by row.

        With the peer-memory exchange (CUDA ranks of one node) the data never goes through a collective: every rank
        pushes its segment of every row straight to the row's final place in the owner's window (``_transpose_push``,
        ``sb200_exchange_push_rows``); only the per-rank row counts (8 bytes per row and rank) are all-gathered.

        Returns (row_bounds, p, cols, vals): this rank owns output columns (= rows of A)
        [row_bounds[rank], row_bounds[rank+1]); p is rebased to 0; cols are GLOBAL column ids of A."""
        W, rank, dev = self.world, self.rank, self.device
        tp, tcols, tvals = self.local.transpose_local()
        tp64 = tp.to(torch.int64)
        counts = (tp64[1:] - tp64[:-1]).contiguous()                      # entries of each row in my block
        if W > 1:
            flat = torch.empty(W * self.nrow, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(flat, counts, group=self.group)
            all_counts = flat.view(W, self.nrow)
        else:
            all_counts = counts.view(1, -1)
        total = all_counts.sum(dim=0)                                       # global entries per row
        P = torch.zeros(self.nrow + 1, dtype=torch.int64, device=dev)
        torch.cumsum(total, 0, out=P[1:])
        nnz = int(P[-1])
        # nnz-balanced contiguous row blocks (same rule as the column split, applied to p')
        targets = torch.tensor([(nnz * k) // W for k in range(1, W)], dtype=torch.int64, device=dev)
        inner = torch.searchsorted(P, targets, right=False).clamp_(0, self.nrow).tolist() if W > 1 else []
        rb = [0] + [max(0, int(v)) for v in inner] + [self.nrow]
        for k in range(1, len(rb)):
            rb[k] = max(rb[k], rb[k - 1])
        r0, r1 = rb[rank], rb[rank + 1]
        if self.window is not None and W > 1:
            return self._transpose_push(rb, P, all_counts, tp, tcols, tvals)
        # what I send to rank k: my entries of rows [rb[k], rb[k+1]) — one contiguous slice of my local transpose
        send_off = [int(tp64[rb[k]]) for k in range(W + 1)]
        send_splits = [send_off[k + 1] - send_off[k] for k in range(W)]
        my_counts = all_counts[:, r0:r1]                                    # [W, R]: segment lengths per (source, my row)
        recv_splits = [int(v) for v in my_counts.sum(dim=1).tolist()]
        gcols = (tcols.to(torch.int64) + self.c0).to(torch.int32)           # global column ids
        n_recv = sum(recv_splits)
        rcols = torch.empty(n_recv, dtype=torch.int32, device=dev)
        rvals = torch.empty(n_recv, dtype=torch.float64, device=dev)
        if W > 1:
            dist.all_to_all_single(rcols, gcols, recv_splits, send_splits, group=self.group)
            dist.all_to_all_single(rvals, tvals.contiguous(), recv_splits, send_splits, group=self.group)
        else:
            rcols.copy_(gcols)
            rvals.copy_(tvals)
        # interleave: destination order is (row, source); the received order is (source, row)
        R = r1 - r0
        p_own = torch.zeros(R + 1, dtype=torch.int64, device=dev)
        torch.cumsum(total[r0:r1], 0, out=p_own[1:])
        if n_recv > 0 and W > 1:
            seg_len = my_counts.t().reshape(-1)                              # [(row, source)]
            recv_base = torch.tensor([0] + recv_splits[:-1], dtype=torch.int64, device=dev).cumsum(0)
            within = torch.cumsum(my_counts, dim=1) - my_counts              # start of row r inside source j's slice
            seg_src = (recv_base.view(-1, 1) + within).t().reshape(-1)       # where each (row, source) segment sits in recv
            seg_dst = torch.cumsum(seg_len, 0) - seg_len                     # and where it goes
            src_index = torch.repeat_interleave(seg_src - seg_dst, seg_len) + torch.arange(n_recv, device=dev)
            rcols, rvals = rcols[src_index], rvals[src_index]
        return rb, p_own.to(torch.int32), rcols, rvals

    def _transpose_push(self, rb, P, all_counts, tp, tcols, tvals):
        """The exchange step of the sharded transpose as P2P stores (sb200_exchange_push_rows): output row r is the
        concatenation, in rank (= column) order, of the ranks' segments of row r; my segment of row r starts
        (P[r] - P[first row of the owner's block]) + sum of the counts of the ranks before me, inside the owner's output."""
        W, rank, dev = self.world, self.rank, self.device
        lib = self.window._lib
        n_out = [int(P[rb[q + 1]] - P[rb[q]]) for q in range(W)]
        align = lambda b: (b + 255) & ~255  # noqa: E731
        # one window per rank for its block of the result: [int32 column ids | float64 values] behind the header
        win = PeerWindow(dev, rank, W, align(4 * n_out[rank]) + 8 * n_out[rank] + 512, self.group)
        try:
            data0 = win._cursor
            cols_off = [data0] * W
            vals_off = [data0 + align(4 * n_out[q]) for q in range(W)]
            before_me = all_counts[:rank].sum(dim=0) if rank > 0 else torch.zeros(self.nrow, dtype=torch.int64, device=dev)
            block_start = torch.empty(self.nrow, dtype=torch.int64, device=dev)
            for q in range(W):
                block_start[rb[q]:rb[q + 1]] = P[rb[q]]
            dst_off = (P[:-1] - block_start + before_me).contiguous()
            rb_h = np.asarray(rb, np.int32)
            co_h, vo_h = np.asarray(cols_off, np.int64), np.asarray(vals_off, np.int64)
            tp32 = tp.to(torch.int32).contiguous()
            st = torch.cuda.current_stream(dev)
            lib.check(lib.lib().sb200_exchange_push_rows(
                win._h, C.c_void_p(st.cuda_stream), C.c_void_p(tp32.data_ptr()), C.c_void_p(tcols.data_ptr() if tcols.numel() else 0),
                C.c_void_p(tvals.data_ptr() if tvals.numel() else 0), C.c_void_p(dst_off.data_ptr()), self.nrow, self.c0,
                C.c_void_p(rb_h.ctypes.data), C.c_void_p(co_h.ctypes.data), C.c_void_p(vo_h.ctypes.data)))
            n = n_out[rank]
            if n > 0:  # the kernel's closing barrier (in stream order) says every rank's segments have landed here
                rcols = torch.as_tensor(_CudaView(win.base + cols_off[rank], n, "<i4"), device=dev).clone()
                rvals = torch.as_tensor(_CudaView(win.base + vals_off[rank], n, "<f8"), device=dev).clone()
            else:
                rcols = torch.empty(0, dtype=torch.int32, device=dev)
                rvals = torch.empty(0, dtype=torch.float64, device=dev)
            win.status()
        finally:
            dist.barrier(group=self.group)  # nobody unmaps while a peer may still write
            win.close()
        r0, r1 = rb[rank], rb[rank + 1]
        p_own = (P[r0:r1 + 1] - P[r0]).to(torch.int32)
        return rb, p_own, rcols, rvals
