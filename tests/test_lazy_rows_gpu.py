"""SB200_LAZY_ROWS (sparse_b200.h): a per-call mirror leaves the row indices on the host until an op reads them.
columnSums / colSums / colMeans never read `i` in the reference (src/example.cpp:28-30, RcppSparse.h:133-135), so they
work without it — even when `i` is garbage, as in the reference; every other op brings `i` over (and checks it) first."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from rcppsparse_b200 import DeviceMatrix, Matrix, SparseB200Error, _lib, columnSums, synth

pytestmark = pytest.mark.gpu


def _case():
    spec = synth.powerlaw_spec(30_000, 4_000, 120.0, 77)
    i, p, x = synth.generate_host(spec)
    return spec, i, p, x


def test_column_sweeps_do_not_need_the_rows_and_everything_else_fetches_them():
    spec, i, p, x = _case()
    a = (i, p, x, spec.nrow, spec.ncol)
    chk = oracle.best()
    v_c, v_r = synth.dense_vector(1, spec.ncol), synth.dense_vector(2, spec.nrow)
    with DeviceMatrix.from_host(i, p, x, spec.nrow, spec.ncol, lazy_rows=True) as M:
        oracle.assert_within("colSums", M.col_sums(), chk.colSums(*a), *a)
        oracle.assert_within("colMeans", M.col_means(), chk.colMeans(*a), *a)
        oracle.assert_within("rowSums", M.row_sums(), chk.rowSums(*a), *a)  # first reader of `i`
        oracle.assert_within("spmv", M.spmv(v_c), chk.spmv(*a, v_c), *a, v=v_c)
        oracle.assert_within("spmv_t", M.spmv_t(v_r), chk.spmv_t(*a, v_r), *a, v=v_r)
    for first in ("transpose", "spmv_t", "rowMeans", "download"):
        with DeviceMatrix.from_host(i, p, x, spec.nrow, spec.ncol, lazy_rows=True) as M:
            if first == "transpose":
                ti, tp, tx = M.transpose_host()
                wi, wp, wx = chk.transpose(*a)
                assert np.array_equal(tp, wp) and np.array_equal(ti, wi) and np.array_equal(tx.view(np.uint64), wx.view(np.uint64))
            elif first == "spmv_t":
                oracle.assert_within("spmv_t", M.spmv_t(v_r), chk.spmv_t(*a, v_r), *a, v=v_r)
            elif first == "rowMeans":
                oracle.assert_within("rowMeans", M.row_means(), chk.rowMeans(*a), *a)
            else:
                di, dp, dx = M.download_columns()
                assert np.array_equal(di, i) and np.array_equal(dp, p)


def test_garbage_rows_only_hurt_the_ops_that_read_them():
    spec, i, p, x = _case()
    bad = i.copy()
    bad[len(bad) // 2] = spec.nrow + 5  # the reference: index_out_of_bounds from rowSums (RcppSparse.h:142), colSums unaffected
    a = (i, p, x, spec.nrow, spec.ncol)
    chk = oracle.best()
    with DeviceMatrix.from_host(bad, p, x, spec.nrow, spec.ncol, lazy_rows=True) as M:
        oracle.assert_within("colSums", M.col_sums(), chk.colSums(*a), *a)
        for _ in range(2):  # fails the same way every time
            with pytest.raises(SparseB200Error) as e:
                M.row_sums()
            assert e.value.code == _lib.E_STRUCTURE
        oracle.assert_within("colMeans", M.col_means(), chk.colMeans(*a), *a)
    with pytest.raises(SparseB200Error):  # without the flag the check happens at create
        DeviceMatrix.from_host(bad, p, x, spec.nrow, spec.ncol)


def test_the_python_mirror_of_the_dropin_uses_it_for_per_call_mirrors():
    spec, i, p, x = _case()
    a = (i, p, x, spec.nrow, spec.ncol)
    chk = oracle.best()
    A = Matrix(x, i, p, np.array([spec.nrow, spec.ncol], np.int32))
    oracle.assert_within("columnSums", columnSums(A), chk.columnSums(*a), *a)
    oracle.assert_within("rowSums", A.rowSums(), chk.rowSums(*a), *a)
    oracle.assert_within("colMeans", A.colMeans(), chk.colMeans(*a), *a)
    T = A.transpose()
    wi, wp, wx = chk.transpose(*a)
    assert np.array_equal(T.p, wp) and np.array_equal(T.i, wi) and np.array_equal(T.x.view(np.uint64), wx.view(np.uint64))
    A.release()
