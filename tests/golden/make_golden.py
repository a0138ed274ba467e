#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the REFERENCE's own compiled code (oracle/_ref).

Run in the build container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden.py

Every output array in the fixtures is produced by liboracle_ref.so, i.e. by the loops of
/root/reference/inst/include/RcppSparse.h:131-156,375-385 and src/example.cpp:26-32 compiled
unmodified (transpose()'s R callee and the two SpMV sweeps are the labelled restatements in
oracle/ref_shim.cpp — the reference holds no code for them).  The reference ships no tests
or golden vectors of its own; the only literal matrix in the tree is
the 5x5 of vignettes/Documentation.Rmd:213-216, which is case "vignette_5x5".
"""
import os
import zlib
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from rcppsparse_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def from_dense(a):
    a = np.asarray(a, dtype=np.float64)
    nrow, ncol = a.shape
    i, x, p = [], [], [0]
    for c in range(ncol):
        nz = np.nonzero(a[:, c])[0]
        i.extend(nz.tolist())
        x.extend(a[nz, c].tolist())
        p.append(len(i))
    return np.array(i, np.int32), np.array(p, np.int32), np.array(x, np.float64), nrow, ncol


def cases():
    rng = np.random.default_rng(20261018)
    out = {}
    # the only literal fixture in the reference tree (vignettes/Documentation.Rmd:213-216)
    out["vignette_5x5"] = (np.array([0, 2, 0, 1, 1], np.int32), np.array([0, 0, 1, 2, 4, 5], np.int32),
                           np.array([0.41, 0.35, 0.84, 0.37, 0.26]), 5, 5)
    out["all_empty_7x4"] = (np.zeros(0, np.int32), np.zeros(5, np.int32), np.zeros(0), 7, 4)
    out["ncol0_5x0"] = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0), 5, 0)
    out["nrow0_0x3"] = (np.zeros(0, np.int32), np.zeros(4, np.int32), np.zeros(0), 0, 3)
    out["one_dense_column_64x1"] = (np.arange(64, dtype=np.int32), np.array([0, 64], np.int32),
                                    rng.standard_normal(64), 64, 1)
    out["one_sparse_column_1000x1"] = (np.array([3, 500, 999], np.int32), np.array([0, 3], np.int32),
                                       np.array([1.5, -2.25, 1e-3]), 1000, 1)
    out["one_row_1x9"] = from_dense(np.array([[1.0, 0, 2.0, 0, 0, -3.0, 4.0, 0, 0.5]]))
    a = rng.standard_normal((17, 13))
    out["all_dense_17x13"] = from_dense(a)
    # stored explicit zeros, NaN, +-Inf are ordinary values (SURVEY.md 8a "Semantics to preserve")
    i, p, x, nr, nc = from_dense(rng.standard_normal((12, 9)) * (rng.random((12, 9)) < 0.4))
    x = x.copy()
    x[1] = 0.0
    x[5] = np.nan
    x[9] = np.inf
    x[12] = -np.inf
    x[20] = -0.0
    out["specials_12x9"] = (i, p, x, nr, nc)
    # a full-height column between runs of empty columns; first and last columns empty
    nrow = 300
    cols = [[], [], list(range(nrow)), [], [], [], [7], [], list(range(0, nrow, 3)), [], []]
    ii = np.array([r for c in cols for r in c], np.int32)
    pp = np.cumsum([0] + [len(c) for c in cols]).astype(np.int32)
    out["max_column_between_empties_300x11"] = (ii, pp, rng.standard_normal(ii.shape[0]), nrow, len(cols))
    # large magnitudes of mixed sign: cancellation stresses the 1e-12*sum|a| criterion
    i, p, x, nr, nc = from_dense(rng.standard_normal((64, 40)) * (rng.random((64, 40)) < 0.5))
    out["cancellation_64x40"] = (i, p, x * 10.0 ** rng.integers(-8, 9, x.shape[0]), nr, nc)
    # seeded synthetic matrices from the bench recipes, scaled down
    for tag, spec in (("synth_uniform_2000x300", synth.uniform_spec(2000, 300, 0.02, 1001)),
                      ("synth_powerlaw_3000x400", synth.powerlaw_spec(3000, 400, 40.0, 1004)),
                      ("synth_banded_1500x500", synth.powerlaw_spec(1500, 500, 60.0, 1003, row_levels=6)),
                      ("synth_wide_50x4000", synth.powerlaw_spec(50, 4000, 3.0, 1005, empty_permille=300))):
        i, p, x = synth.generate_host(spec)
        out[tag] = (i, p, x, spec.nrow, spec.ncol)
    return out


def main():
    ref = oracle.Ref()
    for name, (i, p, x, nrow, ncol) in cases().items():
        seed = zlib.crc32(name.encode()) % 1000 + 7
        v_col = synth.dense_vector(seed, ncol)
        v_row = synth.dense_vector(seed + 100, nrow)
        ti, tp, tx = ref.transpose(i, p, x, nrow, ncol)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            i=i, p=p, x=x, dim=np.array([nrow, ncol], np.int32), v_col=v_col, v_row=v_row,
            columnSums=ref.columnSums(i, p, x, nrow, ncol), colSums=ref.colSums(i, p, x, nrow, ncol),
            rowSums=ref.rowSums(i, p, x, nrow, ncol), colMeans=ref.colMeans(i, p, x, nrow, ncol),
            rowMeans=ref.rowMeans(i, p, x, nrow, ncol), t_i=ti, t_p=tp, t_x=tx,
            spmv=ref.spmv(i, p, x, nrow, ncol, v_col), spmv_t=ref.spmv_t(i, p, x, nrow, ncol, v_row))
        print(f"{name}: {nrow}x{ncol} nnz={x.shape[0]}")


if __name__ == "__main__":
    main()
