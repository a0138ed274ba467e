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
    L.dropin_crossprod.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, _f64]
    L.dropin_transpose.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, _i32, _i32, _f64, _i32]
    L.dropin_alias_semantics.argtypes = [_i32, _i32, _f64, C.c_int, C.c_int, C.c_int64, _f64, _f64]
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
def test_dropin_copies_share_the_mirror_and_refresh(shim):
    i, p = np.array([0, 1, 0], np.int32), np.array([0, 2, 3], np.int32)
    x = np.array([1.0, 2.0, 3.0])
    before, after = np.empty(2), np.empty(2)
    assert shim.dropin_alias_semantics(i, p, x, 2, 2, 3, before, after) == 0, shim.dropin_last_error()
    assert before.tolist() == [3.0, 3.0] and after.tolist() == [1003.0, 3.0]


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
