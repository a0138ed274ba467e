"""The drop-in C++ header (include/RcppSparse.h): user-level code written against RcppSparse::Matrix,
compiled with the Rcpp stand-in and linked to libsparse_b200 (tests/dropin/dropin_shim.cpp)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle

_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def shim():
    import importlib.util
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin", "build_shim.py")
    spec = importlib.util.spec_from_file_location("build_shim", here)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    L = C.CDLL(mod.build())
    L.dropin_last_error.restype = C.c_char_p
    L.dropin_reduce.argtypes = [C.c_int, _i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, _f64]
    L.dropin_spmv.argtypes = [C.c_int, _i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, _f64, _f64]
    _u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    _ip = C.POINTER(C.c_int)
    L.dropin_dense_parts.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, _i32, C.c_int, _i32, C.c_int,
                                     _f64, _f64, _f64, _f64, _f64, _f64, _f64]
    L.dropin_range_cursors.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, C.c_int, _u32, C.c_int, _i32, _f64, _ip,
                                       _i32, _f64, _ip, _u32, _ip, _u32, _ip]
    L.dropin_row_cursor.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, C.c_int, _i32, _f64, _ip, _ip]
    L.dropin_crossprod.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, _f64]
    L.dropin_transpose.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, _i32, _i32, _f64, _i32]
    L.dropin_alias_semantics.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, _f64, _f64, _f64, _f64]
    L.dropin_repointed_members.argtypes = [_i32, _i32, _f64, _f64, C.c_int, C.c_int, C.c_int64, _f64, _f64]
    return L


def test_header_compiles_and_rejects_bad_s4(shim):
    """CPU-side: the header builds against the stand-in; a dgCMatrix without all four slots throws
    std::invalid_argument exactly like the reference (RcppSparse.h:35-36)."""
    assert shim.dropin_missing_slot() == 1
    assert b"Cannot construct RcppSparse::Matrix from this S4 object" in shim.dropin_last_error()


def test_without_gpu_methods_throw(shim):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    i, p, x = np.array([0], np.int32), np.array([0, 1], np.int32), np.array([2.0])
    out = np.zeros(1)
    assert shim.dropin_reduce(1, i, p, x, 1, 1, 1, out) == 2  # std::runtime_error, never a silent CPU path
    assert b"no CUDA device" in shim.dropin_last_error()


@pytest.mark.gpu
def test_dropin_class_matches_golden(shim, golden):
    g = golden
    i, p, x, nrow, ncol = g["i"], g["p"], g["x"], g["nrow"], g["ncol"]
    nnz = x.shape[0]
    args = (i, p, x, nrow, ncol)
    for op, name, n in ((0, "columnSums", ncol), (1, "colSums", ncol), (2, "rowSums", nrow), (3, "colMeans", ncol),
                        (4, "rowMeans", nrow)):
        out = np.empty(n)
        assert shim.dropin_reduce(op, i, p, x, nrow, ncol, nnz, out) == 0, shim.dropin_last_error()
        oracle.assert_within(name, out, g[name], *args)
    # a user's own InnerIterator loop still runs on the host, bit-identical to the reference example
    out = np.empty(ncol)
    assert shim.dropin_reduce(5, i, p, x, nrow, ncol, nnz, out) == 0
    assert np.array_equal(out.view(np.uint64), g["columnSums"].view(np.uint64))
    y = np.empty(nrow)
    assert shim.dropin_spmv(0, i, p, x, nrow, ncol, nnz, g["v_col"], y) == 0, shim.dropin_last_error()
    oracle.assert_within("spmv", y, g["spmv"], *args, v=g["v_col"])
    y = np.empty(ncol)
    assert shim.dropin_spmv(1, i, p, x, nrow, ncol, nnz, g["v_row"], y) == 0, shim.dropin_last_error()
    oracle.assert_within("spmv_t", y, g["spmv_t"], *args, v=g["v_row"])
    if ncol <= 2000:
        cp = np.empty((ncol, ncol))
        assert shim.dropin_crossprod(i, p, x, nrow, ncol, nnz, cp.reshape(-1)) == 0, shim.dropin_last_error()
        oracle.assert_within("crossprod", cp, oracle.best().crossprod(*args), *args)
        assert np.array_equal(cp.view(np.uint64), cp.T.view(np.uint64)), "crossprod must be exactly symmetric"
    tp, ti, tx, td = np.empty(nrow + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz), np.empty(2, np.int32)
    assert shim.dropin_transpose(i, p, x, nrow, ncol, nnz, tp, ti, tx, td) == 0, shim.dropin_last_error()
    assert td.tolist() == [ncol, nrow]
    assert np.array_equal(tp, g["t_p"]) and np.array_equal(ti, g["t_i"])
    assert np.array_equal(tx.view(np.uint64), g["t_x"].view(np.uint64))


@pytest.mark.gpu
def test_dropin_reads_live_by_default_and_resident_mirrors_need_refresh(shim):
    """The reference's methods loop over the R vectors on every call (RcppSparse.h:131-156) and the vignette edits
    them in place (Documentation.Rmd:325-347): by default an edit is seen by the next call.  A resident Matrix keeps
    its device mirror (shared by copies) and sees the edit after refresh()."""
    i, p = np.array([0, 1, 0], np.int32), np.array([0, 2, 3], np.int32)
    x = np.array([1.0, 2.0, 3.0])
    before, after, stale, refreshed = np.empty(2), np.empty(2), np.empty(2), np.empty(2)
    assert shim.dropin_alias_semantics(i, p, x, 2, 2, 3, before, after, stale, refreshed) == 0, shim.dropin_last_error()
    assert before.tolist() == [3.0, 3.0] and after.tolist() == [1003.0, 3.0]
    assert stale.tolist() == [1003.0, 3.0] and refreshed.tolist() == [2003.0, 3.0]


@pytest.mark.gpu
def test_dropin_header_on_two_gpus(shim, monkeypatch):
    """SB200_GPUS=2: the same user code, the sweeps behind sb200_sharded_* on two GPUs of this process."""
    from rcppsparse_b200 import _lib, synth

    if _lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("SB200_GPUS", "2")
    spec = synth.config("C3", 0.002)
    i, p, x = synth.generate_host(spec)
    args = (i, p, x, spec.nrow, spec.ncol)
    chk = oracle.best()
    for op, name in ((1, "colSums"), (2, "rowSums"), (3, "colMeans"), (4, "rowMeans")):
        out = np.empty(spec.ncol if op in (1, 3) else spec.nrow)
        assert shim.dropin_reduce(op, i, p, x, spec.nrow, spec.ncol, x.shape[0], out) == 0, shim.dropin_last_error()
        oracle.assert_within(name, out, getattr(chk, name)(*args), *args)
    v = synth.dense_vector(3, spec.ncol)
    y = np.empty(spec.nrow)
    assert shim.dropin_spmv(0, i, p, x, spec.nrow, spec.ncol, x.shape[0], v, y) == 0, shim.dropin_last_error()
    oracle.assert_within("spmv", y, chk.spmv(*args, v), *args, v=v)
    # Matrix::transpose() on two GPUs (sb200_sharded_transpose): bit for bit
    tp, ti, tx, td = np.empty(spec.nrow + 1, np.int32), np.empty(x.shape[0], np.int32), np.empty(x.shape[0]), np.empty(2, np.int32)
    assert shim.dropin_transpose(i, p, x, spec.nrow, spec.ncol, x.shape[0], tp, ti, tx, td) == 0, shim.dropin_last_error()
    wi, wp, wx = chk.transpose(*args)
    assert td.tolist() == [spec.ncol, spec.nrow] and np.array_equal(tp, wp) and np.array_equal(ti, wi)
    assert np.array_equal(tx.view(np.uint64), wx.view(np.uint64))


@pytest.mark.gpu
def test_dropin_repointed_members_rebuild_the_mirror(shim):
    i, p = np.array([0, 1, 0], np.int32), np.array([0, 2, 3], np.int32)
    x1, x2 = np.array([1.0, 2.0, 3.0]), np.array([10.0, 20.0, 30.0])
    s1, s2 = np.empty(2), np.empty(2)
    assert shim.dropin_repointed_members(i, p, x1, x2, 2, 2, 3, s1, s2) == 0, shim.dropin_last_error()
    assert s1.tolist() == [3.0, 3.0] and s2.tolist() == [30.0, 30.0]


@pytest.mark.gpu
def test_dropin_corrupt_index_is_an_exception(shim):
    i, p, x = np.array([0, 7], np.int32), np.array([0, 2], np.int32), np.ones(2)
    out = np.empty(3)
    assert shim.dropin_reduce(2, i, p, x, 3, 1, 2, out) == 1  # invalid_argument; reference: index_out_of_bounds
    assert b"row index outside" in shim.dropin_last_error()


def test_r_package_glue_type_checks():
    """r-package/src/exports.cpp (the exported R entry points and the persistent external-pointer handle) against
    the drop-in header; R and Rcpp are absent, so the test-only stand-in provides the Rcpp names."""
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(root, "include"),
                        "-I", os.path.join(root, "oracle", "stub"),
                        "-include", os.path.join(root, "tests", "dropin", "rcpp_xptr_stub.h"),
                        os.path.join(root, "r-package", "src", "exports.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def _small_matrix(seed, nrow=23, ncol=17, density=0.3, symmetric=False):
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    dense = np.where(rng.random((nrow, ncol)) < density, np.round(rng.standard_normal((nrow, ncol)), 2), 0.0)
    dense[:, 3] = 0.0  # an empty column
    dense[5, :] = 0.0  # an empty row
    if symmetric:
        dense = np.triu(dense[:ncol, :ncol]) + np.triu(dense[:ncol, :ncol], 1).T
    A = sp.csc_matrix(dense)
    A.sort_indices()
    return dense, A.indices.astype(np.int32), A.indptr.astype(np.int32), A.data.astype(np.float64)


def test_dropin_host_side_dense_parts(shim):
    """Dense copies of rows, columns and blocks (reference RcppSparse.h:75-128) are host-side lookups: no GPU."""
    dense, i, p, x = _small_matrix(1)
    nrow, ncol = dense.shape
    rows, cols = np.array([0, 5, 22, 7, 7], np.int32), np.array([3, 0, 16, 9], np.int32)
    r, c = 11, 9
    row_out, col_out = np.empty(ncol), np.empty(nrow)
    block, rowsel, colsel = np.empty(len(rows) * len(cols)), np.empty(len(cols)), np.empty(len(rows))
    cols_out, rows_out = np.empty(nrow * len(cols)), np.empty(len(rows) * ncol)
    assert shim.dropin_dense_parts(i, p, x, nrow, ncol, len(x), r, c, rows, len(rows), cols, len(cols), row_out, col_out, block,
                                   rowsel, colsel, cols_out, rows_out) == 0, shim.dropin_last_error()
    assert np.array_equal(row_out, dense[r, :]) and np.array_equal(col_out, dense[:, c])
    assert np.array_equal(rowsel, dense[r, cols]) and np.array_equal(colsel, dense[rows, c])
    assert np.array_equal(block.reshape(len(cols), len(rows)).T, dense[np.ix_(rows, cols)])
    assert np.array_equal(cols_out.reshape(len(cols), nrow).T, dense[:, cols])
    assert np.array_equal(rows_out.reshape(ncol, len(rows)).T, dense[rows, :])


@pytest.mark.parametrize("col", [0, 3, 9, 16])
def test_dropin_range_cursors(shim, col):
    """InnerIteratorInRange / NotInRange, InnerIndices, emptyInnerIndices (reference :196-216, :238-321)."""
    dense, i, p, x = _small_matrix(2)
    nrow, ncol = dense.shape
    for s in (np.array([], np.uint32), np.arange(nrow, dtype=np.uint32), np.array([0, 2, 5, 6, 7, 11, 19, 22], np.uint32)):
        in_rows, in_vals, out_rows, out_vals = (np.empty(nrow, np.int32), np.empty(nrow), np.empty(nrow, np.int32), np.empty(nrow))
        inner, empty = np.empty(nrow, np.uint32), np.empty(nrow, np.uint32)
        n_in, n_out, n_inner, n_empty = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        s_arg = s if len(s) else np.zeros(1, np.uint32)
        assert shim.dropin_range_cursors(i, p, x, nrow, ncol, len(x), col, s_arg, len(s), in_rows, in_vals, C.byref(n_in), out_rows,
                                         out_vals, C.byref(n_out), inner, C.byref(n_inner), empty, C.byref(n_empty)) == 0
        stored = np.nonzero(dense[:, col])[0]
        want_in = np.array([r for r in stored if r in set(s.tolist())], np.int64)
        want_out = np.array([r for r in stored if r not in set(s.tolist())], np.int64)
        assert in_rows[:n_in.value].tolist() == want_in.tolist() and np.array_equal(in_vals[:n_in.value], dense[want_in, col])
        assert out_rows[:n_out.value].tolist() == want_out.tolist() and np.array_equal(out_vals[:n_out.value], dense[want_out, col])
        assert inner[:n_inner.value].tolist() == stored.tolist()
        assert empty[:n_empty.value].tolist() == [r for r in range(nrow) if dense[r, col] == 0.0]


def test_dropin_row_cursor_and_symmetry(shim):
    """InnerRowIterator walks a row's stored entries in column order; isAppxSymmetric tests A(r,c) == A(c,r)."""
    dense, i, p, x = _small_matrix(3)
    nrow, ncol = dense.shape
    for row in (0, 5, 11, 22):
        cols, vals, n, sym = np.empty(ncol, np.int32), np.empty(ncol), C.c_int(), C.c_int()
        assert shim.dropin_row_cursor(i, p, x, nrow, ncol, len(x), row, cols, vals, C.byref(n), C.byref(sym)) == 0
        want = np.nonzero(dense[row, :])[0]
        assert cols[:n.value].tolist() == want.tolist() and np.array_equal(vals[:n.value], dense[row, want])
        assert sym.value == 0  # not square
    sd, si, sp_, sx = _small_matrix(4, symmetric=True)
    cols, vals, n, sym = np.empty(sd.shape[1], np.int32), np.empty(sd.shape[1]), C.c_int(), C.c_int()
    assert shim.dropin_row_cursor(si, sp_, sx, sd.shape[0], sd.shape[1], len(sx), 1, cols, vals, C.byref(n), C.byref(sym)) == 0
    assert sym.value == 1
    sx2 = sx.copy()
    sx2[len(sx2) // 2] += 1.0  # one entry off: no longer symmetric (unless it sits on the diagonal)
    k = len(sx2) // 2
    col_of = np.searchsorted(sp_, k, side="right") - 1
    if si[k] != col_of:
        assert shim.dropin_row_cursor(si, sp_, sx2, sd.shape[0], sd.shape[1], len(sx2), 1, cols, vals, C.byref(n), C.byref(sym)) == 0
        assert sym.value == 0
