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


class ShardedMatrix:
    """The reference's Matrix methods (RcppSparse.h:131-156 + the SpMV idiom) over a column-sharded matrix.
    Every method returns the FULL result vector on every rank — as a view of an internal buffer that
    the next call of the same kind (column- or row-indexed) overwrites: clone it to keep it."""

    def __init__(self, local: LocalSweeps, bounds: Sequence[int], rank: int, group=None, device=None):
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
        self._col_full = torch.empty(self.ncol, dtype=torch.float64, device=self.device)
        self._row_full = torch.empty(self.nrow, dtype=torch.float64, device=self.device)

    # ---- column-indexed: disjoint slices + all-gather --------------------------------------------------
    def _gather_columns(self) -> torch.Tensor:
        mine = self._col_full[self.c0:self.c1]
        if self.world == 1:
            return self._col_full
        if self._even:
            dist.all_gather_into_tensor(self._col_full, mine.clone(), group=self.group)
        else:
            parts = [self._col_full[self.bounds[k]:self.bounds[k + 1]] for k in range(self.world)]
            # uneven counts: one broadcast per owner (grouped by the backend)
            for k in range(self.world):
                if parts[k].numel():
                    dist.broadcast(parts[k], src=dist.get_global_rank(self.group, k) if self.group else k,
                                   group=self.group)
        return self._col_full

    def colSums(self) -> torch.Tensor:
        self.local.col_sums(self._col_full[self.c0:self.c1], 0.0)
        return self._gather_columns()

    def colMeans(self) -> torch.Tensor:
        self.local.col_sums(self._col_full[self.c0:self.c1], float(self.nrow))  # nrow is global already
        return self._gather_columns()

    def spmv_t(self, v: torch.Tensor) -> torch.Tensor:
        """A^T v: v[nrow] replicated on every rank."""
        self.local.spmv_t(v, self._col_full[self.c0:self.c1])
        return self._gather_columns()

    # ---- row-indexed: full-length partials + all-reduce ----------------------------------------------------
    def _reduce_rows(self) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(self._row_full, op=dist.ReduceOp.SUM, group=self.group)
        return self._row_full

    def rowSums(self) -> torch.Tensor:
        self.local.row_sums(self._row_full)
        return self._reduce_rows()

    def rowMeans(self) -> torch.Tensor:
        self.local.row_sums(self._row_full)
        out = self._reduce_rows()
        self.local.div(out, float(self.ncol))  # GLOBAL ncol (RcppSparse.h:154 divides by Dim[1])
        return out

    def spmv(self, v: torch.Tensor) -> torch.Tensor:
        """A v: v[ncol] replicated; each rank uses only its own slice."""
        self.local.spmv(v[self.c0:self.c1], self._row_full)
        return self._reduce_rows()

    # ---- transpose: local transposes + ONE exchange step (all-to-all-v) --------------------------------------
    def transpose(self):
        """CSC of A^T, row-sharded.  Every rank transposes its own column block with the
        device kernel; row r of the result is the concatenation, in rank (= column) order, of each rank's
        row-r segment, so order inside a row is ascending source column by construction.  Rows are then
        dealt out in nnz-balanced contiguous blocks: one all-gather of the per-rank row counts, one
        all-to-all-v of (column id, value) pairs, and an index gather that interleaves the received
        per-rank segments row by row.

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
