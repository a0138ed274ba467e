"""CPU tests of the oracle itself (no GPU): the C port and, where present, the compiled
reference must reproduce the committed golden vectors bit for bit; transpose and SpMV —
for which the reference holds no code — are cross-checked against scipy.sparse."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC_OPS = ("columnSums", "colSums", "rowSums", "colMeans", "rowMeans")


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def port():
    return oracle.Port()


def test_vignette_known_answers(port):
    """Known-answer test on the only literal matrix in the reference tree
    (vignettes/Documentation.Rmd:213-216); expected values are SURVEY.md 8(c)'s."""
    i = np.array([0, 2, 0, 1, 1], np.int32)
    p = np.array([0, 0, 1, 2, 4, 5], np.int32)
    x = np.array([0.41, 0.35, 0.84, 0.37, 0.26])
    assert port.colSums(i, p, x, 5, 5).tolist() == [0.0, 0.41, 0.35, 1.21, 0.26]
    assert port.columnSums(i, p, x, 5, 5).tolist() == [0.0, 0.41, 0.35, 1.21, 0.26]
    assert port.rowSums(i, p, x, 5, 5).tolist() == [1.25, 0.63, 0.35, 0.0, 0.0]
    assert port.colMeans(i, p, x, 5, 5).tolist() == [0.0, 0.08199999999999999, 0.069999999999999993,
                                                       0.24199999999999999, 0.052000000000000005]
    assert port.rowMeans(i, p, x, 5, 5).tolist() == [0.25, 0.126, 0.069999999999999993, 0.0, 0.0]
    ti, tp, tx = port.transpose(i, p, x, 5, 5)
    assert tp.tolist() == [0, 2, 4, 5, 5, 5] and ti.tolist() == [1, 3, 3, 4, 2]
    assert tx.tolist() == [0.41, 0.84, 0.37, 0.26, 0.35]
    v = np.arange(1.0, 6.0)
    assert port.spmv(i, p, x, 5, 5, v).tolist() == [4.18, 2.7800000000000002, 1.0499999999999998, 0.0, 0.0]
    assert port.spmv_t(i, p, x, 5, 5, v).tolist() == [0.0, 0.41, 1.0499999999999998, 1.58, 0.52]


def test_port_matches_golden_bitwise(port, golden):
    g = golden
    for op in VEC_OPS:
        got = getattr(port, op)(g["i"], g["p"], g["x"], g["nrow"], g["ncol"])
        assert np.array_equal(bits(got), bits(g[op])), (g["name"], op)
    ti, tp, tx = port.transpose(g["i"], g["p"], g["x"], g["nrow"], g["ncol"])
    assert np.array_equal(ti, g["t_i"]) and np.array_equal(tp, g["t_p"])
    assert np.array_equal(bits(tx), bits(g["t_x"]))
    assert np.array_equal(bits(port.spmv(g["i"], g["p"], g["x"], g["nrow"], g["ncol"], g["v_col"])), bits(g["spmv"]))
    assert np.array_equal(bits(port.spmv_t(g["i"], g["p"], g["x"], g["nrow"], g["ncol"], g["v_row"])),
                          bits(g["spmv_t"]))


@pytest.mark.skipif(not oracle.Ref.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_compiled_reference_matches_golden_bitwise(golden):
    g = golden
    ref = oracle.Ref()
    for op in VEC_OPS:
        got = getattr(ref, op)(g["i"], g["p"], g["x"], g["nrow"], g["ncol"])
        assert np.array_equal(bits(got), bits(g[op])), (g["name"], op)
    ti, tp, tx = ref.transpose(g["i"], g["p"], g["x"], g["nrow"], g["ncol"])
    assert np.array_equal(ti, g["t_i"]) and np.array_equal(tp, g["t_p"]) and np.array_equal(bits(tx), bits(g["t_x"]))


def test_transpose_and_spmv_agree_with_scipy(golden):
    """The pieces the reference delegates (Matrix::t) or lacks (SpMV): independent cross-check."""
    g = golden
    nrow, ncol = g["nrow"], g["ncol"]
    if nrow == 0 or ncol == 0:
        pytest.skip("scipy rejects zero-sized index arrays inconsistently")
    a = sp.csc_matrix((g["x"], g["i"], g["p"]), shape=(nrow, ncol))
    r = a.tocsr()
    r.sort_indices()
    assert np.array_equal(r.indptr.astype(np.int32), g["t_p"])
    assert np.array_equal(r.indices.astype(np.int32), g["t_i"])
    assert np.array_equal(bits(r.data), bits(g["t_x"]))
    if np.isfinite(g["x"]).all():
        oracle.assert_within("spmv", a @ g["v_col"], g["spmv"], g["i"], g["p"], g["x"], nrow, ncol, g["v_col"])
        oracle.assert_within("spmv_t", a.T @ g["v_row"], g["spmv_t"], g["i"], g["p"], g["x"], nrow, ncol, g["v_row"])
        oracle.assert_within("rowSums", np.asarray(a.sum(axis=1)).ravel(), g["rowSums"], g["i"], g["p"], g["x"],
                             nrow, ncol)
        oracle.assert_within("colSums", np.asarray(a.sum(axis=0)).ravel(), g["colSums"], g["i"], g["p"], g["x"],
                             nrow, ncol)


def test_reference_bounds_check_is_an_error_not_ub():
    """RcppSparse.h:142 indexes with Rcpp's checked operator(): a corrupt row index raises."""
    if not oracle.Ref.available():
        pytest.skip("oracle/_ref not built")
    ref = oracle.Ref()
    i = np.array([0, 9], np.int32)  # 9 >= nrow
    p = np.array([0, 2], np.int32)
    with pytest.raises(RuntimeError):
        ref.rowSums(i, p, np.ones(2), 3, 1)


def test_crossprod_port_equals_reference_and_scipy():
    """crossprod (reference RcppSparse.h:158-194): the compiled reference, the C port (same merges, same order) and
    scipy's A.T @ A; the reference's result is exactly symmetric."""
    import scipy.sparse as sp

    from rcppsparse_b200 import synth

    port = oracle.Port()
    for spec in (synth.powerlaw_spec(500, 120, 20.0, 3), synth.uniform_spec(300, 64, 0.2, 4), synth.uniform_spec(50, 1, 0.5, 5)):
        i, p, x = synth.generate_host(spec)
        a = port.crossprod(i, p, x, spec.nrow, spec.ncol)
        want = (sp.csc_matrix((x, i, p), shape=(spec.nrow, spec.ncol)).T @ sp.csc_matrix((x, i, p), shape=(spec.nrow, spec.ncol))).toarray()
        oracle.assert_within("crossprod", a, want, i, p, x, spec.nrow, spec.ncol)
        assert np.array_equal(a, a.T)
        if oracle.Ref.available():
            b = oracle.Ref().crossprod(i, p, x, spec.nrow, spec.ncol)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), "port and reference must agree bit for bit"


def test_crossprod_known_answer_on_the_vignette_matrix():
    """The one literal matrix in the reference tree (vignettes/Documentation.Rmd:213-216), A^T A worked out by hand."""
    x = np.array([0.41, 0.35, 0.84, 0.37, 0.26])
    i = np.array([0, 2, 0, 1, 1], np.int32)
    p = np.array([0, 0, 1, 2, 4, 5], np.int32)
    want = np.zeros((5, 5))
    want[1, 1] = 0.41 * 0.41
    want[1, 3] = want[3, 1] = 0.41 * 0.84
    want[2, 2] = 0.35 * 0.35
    want[3, 3] = 0.84 * 0.84 + 0.37 * 0.37
    want[3, 4] = want[4, 3] = 0.37 * 0.26
    want[4, 4] = 0.26 * 0.26
    for chk in ([oracle.Port()] + ([oracle.Ref()] if oracle.Ref.available() else [])):
        got = chk.crossprod(i, p, x, 5, 5)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), chk.kind


@pytest.mark.parametrize("n_super,chunk", [(1, 1000), (8, 1000), (5, 37), (13, 100000)])
def test_two_level_transpose_spec_is_the_canonical_transpose(n_super, chunk):
    """tools/transpose_two_level_spec.py (the specification of transpose_split.cu): a stable partition into
    row super-bands followed by a stable split inside each is the counting-sort transpose, bit for bit."""
    import importlib.util

    from rcppsparse_b200 import synth

    spec_ = importlib.util.spec_from_file_location("two_level", os.path.join(ROOT, "tools", "transpose_two_level_spec.py"))
    mod = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(mod)
    for spec in (synth.powerlaw_spec(3000, 400, 25.0, 5, row_levels=4), synth.uniform_spec(20000, 150, 0.004, 6)):
        i, p, x = synth.generate_host(spec)
        ri, rp, rx = oracle.Port().transpose(i, p, x, spec.nrow, spec.ncol)
        for shift in (None, 6):  # equal row counts per band, or the device's bands (row >> shift) with its 4096-entry chunks
            ti, tp, tx = mod.transpose_two_level(i, p, x, spec.nrow, spec.ncol, n_super, 4096 if shift else chunk, shift)
            assert np.array_equal(tp, rp) and np.array_equal(ti, ri) and np.array_equal(tx.view(np.uint64), rx.view(np.uint64))
