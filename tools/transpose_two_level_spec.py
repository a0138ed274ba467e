#!/usr/bin/env python
"""Executable specification (numpy) of the two-level stable split used for the transpose of TALL matrices
(rcppsparse_b200/csrc/transpose_split.cu) — not product code.  The device path reproduces these
intermediate arrays with bands of 2^shift rows (`shift=`), chunks of 4096 entries and 16-byte records; its destination
table is `offs` below (exclusive scan over (band major, chunk minor) of per-chunk band counts).

Why: with 2432-row bands a 1M-row matrix has ~450 bands, a band's run per column is 1-2 entries, every entry costs
two scattered partial-sector stores (~20 G requests/s on B200, profiles/r01/microbench_stores.jsonl) and the band
pointers alone are nb x ncol x 4 bytes.  Two levels keep both passes coalesced:

  level 1  stable partition of the entry stream (storage order = column-major) into B super-bands of rows.  Per
           chunk of CH consecutive entries and per super-band: count -> exclusive scan over (super-band, chunk) ->
           every chunk writes its entries of a super-band as ONE contiguous run.  Output per super-band: its entries
           in the original (column) order as (row, col, value) records, plus its own column pointer.
  level 2  inside a super-band (its entries contiguous and in column order) a stable split by row of every chunk of
           its stream, each row's piece appended at the row's cursor.

Stability of both levels makes the result the canonical CSC of A^T, bit for bit (checked below against the
counting-sort oracle).
"""
import numpy as np


def level1_partition(i, p, x, nrow, ncol, n_super, chunk, shift=None):
    """Returns per super-band b: (rows, cols, vals) in original order, exactly as a device pass would write them:
    run offsets come from an exclusive scan over (super-band major, chunk minor) of per-chunk counts."""
    nnz = len(x)
    col_of = np.repeat(np.arange(ncol, dtype=np.int32), np.diff(p))
    if shift is not None:  # the device's bands: row >> shift
        n_super = (nrow + (1 << shift) - 1) >> shift if nrow else 1
        bounds = np.minimum(np.arange(n_super + 1, dtype=np.int64) << shift, nrow)
    else:
        bounds = (np.arange(n_super + 1, dtype=np.int64) * nrow) // n_super  # equal row counts per super-band
    band_of = np.searchsorted(bounds, i, side="right") - 1
    n_chunks = (nnz + chunk - 1) // chunk
    counts = np.zeros((n_super, n_chunks), np.int64)
    for c in range(n_chunks):
        counts[:, c] = np.bincount(band_of[c * chunk:(c + 1) * chunk], minlength=n_super)
    offs = np.concatenate([[0], np.cumsum(counts.reshape(-1))])[:-1].reshape(n_super, n_chunks)
    out_r, out_c, out_v = np.empty(nnz, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    for c in range(n_chunks):
        sl = slice(c * chunk, min((c + 1) * chunk, nnz))
        b = band_of[sl]
        for band in range(n_super):
            sel = np.nonzero(b == band)[0]  # in order: the chunk's run for this super-band keeps storage order
            o = offs[band, c]
            out_r[o:o + len(sel)] = i[sl][sel]
            out_c[o:o + len(sel)] = col_of[sl][sel]
            out_v[o:o + len(sel)] = x[sl][sel]
    starts = np.concatenate([offs[:, 0], [nnz]])
    return bounds, starts, out_r, out_c, out_v


def level2_transpose(bounds, starts, r, c, v, nrow):
    """Stable counting sort by row inside each super-band; concatenation over super-bands is the transpose."""
    tp = np.zeros(nrow + 1, np.int64)
    np.add.at(tp, r.astype(np.int64) + 1, 1)
    tp = np.cumsum(tp)
    ti, tx = np.empty(len(v), np.int32), np.empty(len(v), np.float64)
    for b in range(len(bounds) - 1):
        s, e = starts[b], starts[b + 1]
        order = np.argsort(r[s:e], kind="stable")  # rows ascending, ties in (column) arrival order
        ti[s:e] = c[s:e][order]  # a super-band's rows are contiguous in the output: tp[bounds[b]] == s
        tx[s:e] = v[s:e][order]
        assert tp[bounds[b]] == s
    return ti, tp.astype(np.int32), tx


def transpose_two_level(i, p, x, nrow, ncol, n_super=8, chunk=1000, shift=None):
    bounds, starts, r, c, v = level1_partition(i, p, x, nrow, ncol, n_super, chunk, shift)
    return level2_transpose(bounds, starts, r, c, v, nrow)
