"""The single-process column-sharding layer behind the C ABI (sb200_sharded_*, sparse_b200.h): one handle, n_gpus
devices, host vectors in and out — what the drop-in header drives when SB200_GPUS > 1.  Against the oracle on the
whole matrix.  n_gpus = 1 runs anywhere; more needs that many GPUs with peer access (skipped otherwise)."""
import numpy as np
import pytest

from oracle import oracle
from rcppsparse_b200 import ShardedHostMatrix, SparseB200Error, _lib, synth

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = {
    "C2_scaled": lambda: synth.config("C2", 0.03),
    "C3_scaled": lambda: synth.config("C3", 0.003),
    "powerlaw": lambda: synth.powerlaw_spec(40_000, 9_001, 150.0, 31),
    "fewer_columns_than_gpus": lambda: synth.uniform_spec(5000, 1, 0.3, 5),
}


def gpu_counts():
    n = _lib.device_count()
    return [g for g in (1, 2, 4, 8) if g <= max(n, 1)]


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
@pytest.mark.parametrize("case", sorted(CASES))
def test_sharded_handle_matches_the_reference(case, n_gpus):
    if n_gpus > _lib.device_count():
        pytest.skip(f"needs {n_gpus} GPUs")
    spec = CASES[case]()
    i, p, x = synth.generate_host(spec)
    a = (i, p, x, spec.nrow, spec.ncol)
    chk = oracle.best()
    v_c, v_r = synth.dense_vector(1, spec.ncol), synth.dense_vector(2, spec.nrow)
    with ShardedHostMatrix(i, p, x, spec.nrow, spec.ncol, n_gpus) as S:
        b = S.bounds()
        assert b[0] == 0 and b[-1] == spec.ncol and all(b[k] <= b[k + 1] for k in range(n_gpus))
        for rep in range(10):  # the ninth call of each kind switches every block to its cached layout
            oracle.assert_within("colSums", S.col_sums(), chk.colSums(*a), *a, tol=TOL)
            oracle.assert_within("colMeans", S.col_means(), chk.colMeans(*a), *a, tol=TOL)
            oracle.assert_within("rowSums", S.row_sums(), chk.rowSums(*a), *a, tol=TOL)
            oracle.assert_within("rowMeans", S.row_means(), chk.rowMeans(*a), *a, tol=TOL)
            oracle.assert_within("spmv", S.spmv(v_c), chk.spmv(*a, v_c), *a, v=v_c, tol=TOL)
            oracle.assert_within("spmv_t", S.spmv_t(v_r), chk.spmv_t(*a, v_r), *a, v=v_r, tol=TOL)
        r1, r2 = S.row_sums(), S.row_sums()
        assert np.array_equal(r1.view(np.uint64), r2.view(np.uint64)), "rank-ordered reduction of deterministic partials"


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
@pytest.mark.parametrize("case", sorted(CASES) + ["C2_5pct"])
def test_sharded_transpose_bit_exact(case, n_gpus):
    """sb200_sharded_transpose: local transposes, row pieces pushed over peer memory in block (= column) order, every
    device copies its rows home — the canonical CSC of A^T bit for bit, twice (cached plans the second time)."""
    if n_gpus > _lib.device_count():
        pytest.skip(f"needs {n_gpus} GPUs")
    spec = synth.config("C2", 0.05) if case == "C2_5pct" else CASES[case]()
    i, p, x = synth.generate_host(spec)
    wi, wp, wx = oracle.best().transpose(i, p, x, spec.nrow, spec.ncol)
    with ShardedHostMatrix(i, p, x, spec.nrow, spec.ncol, n_gpus) as S:
        for _ in range(2):
            ti, tp, tx = S.transpose_host()
            assert np.array_equal(tp, wp) and np.array_equal(ti, wi) and np.array_equal(tx.view(np.uint64), wx.view(np.uint64))


def test_sharded_transpose_golden_edges(golden):
    g = golden
    for n_gpus in gpu_counts()[:2]:
        with ShardedHostMatrix(g["i"], g["p"], g["x"], g["nrow"], g["ncol"], n_gpus) as S:
            ti, tp, tx = S.transpose_host()
        assert np.array_equal(tp, g["t_p"]) and np.array_equal(ti, g["t_i"]) and np.array_equal(
            np.asarray(tx).view(np.uint64), np.asarray(g["t_x"], np.float64).view(np.uint64))


def test_sharded_handle_golden_edges(golden):
    g = golden
    a = (g["i"], g["p"], g["x"], g["nrow"], g["ncol"])
    for n_gpus in gpu_counts()[:2]:
        with ShardedHostMatrix(*a, n_gpus) as S:
            oracle.assert_within("colSums", S.col_sums(), g["colSums"], *a, tol=TOL)
            oracle.assert_within("rowSums", S.row_sums(), g["rowSums"], *a, tol=TOL)
            oracle.assert_within("colMeans", S.col_means(), g["colMeans"], *a, tol=TOL)
            oracle.assert_within("rowMeans", S.row_means(), g["rowMeans"], *a, tol=TOL)
            oracle.assert_within("spmv", S.spmv(g["v_col"]), g["spmv"], *a, v=g["v_col"], tol=TOL)
            oracle.assert_within("spmv_t", S.spmv_t(g["v_row"]), g["spmv_t"], *a, v=g["v_row"], tol=TOL)


def test_sharded_create_reports_errors():
    spec = synth.config("C1", 0.01)
    i, p, x = synth.generate_host(spec)
    with pytest.raises(SparseB200Error):
        ShardedHostMatrix(i, p, x, spec.nrow, spec.ncol, 0)
    with pytest.raises(SparseB200Error):
        ShardedHostMatrix(i, p, x, spec.nrow, spec.ncol, 1, devices=[99])
    bad = i.copy()
    bad[3] = spec.nrow + 7  # a wild row index is an error from the block that holds it, not undefined behaviour
    with pytest.raises(SparseB200Error):
        ShardedHostMatrix(bad, p, x, spec.nrow, spec.ncol, 1)
