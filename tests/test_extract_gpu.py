"""SURVEY.md 8(f) N4 as device ops: every column's InnerIteratorInRange / InnerIteratorNotInRange sweep at once, and
dense extraction of rows / columns / blocks, against a host restatement of the reference's cursors
(RcppSparse.h:238-321: an entry of the column is visited iff its row is / is not in the index set) and dense numpy."""
import numpy as np
import pytest

from rcppsparse_b200 import DeviceMatrix, synth

pytestmark = pytest.mark.gpu


def masked_col_sums_host(i, p, x, ncol, rows, negate):
    """The user loop `for (InnerIteratorInRange it(A, col, s); it; ++it) sum += it.value()` for every column, serial."""
    sel = np.isin(i, rows)
    if negate:
        sel = ~sel
    out = np.zeros(ncol)
    for c in range(ncol):
        acc = 0.0
        for k in range(p[c], p[c + 1]):
            if sel[k]:
                acc += x[k]
        out[c] = acc
    return out


@pytest.mark.parametrize("negate", [False, True])
@pytest.mark.parametrize("case", ["uniform", "powerlaw_rows", "tiny_columns"])
def test_masked_column_sums_match_the_range_cursors(case, negate):
    spec = {"uniform": synth.uniform_spec(5000, 700, 0.02, 3), "powerlaw_rows": synth.powerlaw_spec(3000, 500, 60.0, 4, row_levels=5),
            "tiny_columns": synth.powerlaw_spec(900, 3000, 2.0, 5, empty_permille=300)}[case]
    i, p, x = synth.generate_host(spec)
    rng = np.random.default_rng(17)
    for frac in (0.0, 0.1, 0.5, 1.0):
        rows = np.sort(rng.choice(spec.nrow, int(frac * spec.nrow), replace=False)).astype(np.int32)
        want = masked_col_sums_host(i, p, x, spec.ncol, rows, negate)
        with DeviceMatrix.from_host(i, p, x, spec.nrow, spec.ncol) as D:
            got = D.col_sums_in_rows(rows, negate)
            feed = np.bincount(np.repeat(np.arange(spec.ncol), np.diff(p)), weights=np.abs(x), minlength=spec.ncol)
            assert np.all(np.abs(got - want) <= 1e-12 * feed), (case, frac, negate)
            # unsorted, duplicated and out-of-range indices select the same rows
            noisy = np.concatenate([rows[::-1], rows[:5], np.array([-3, spec.nrow + 9], np.int32)]).astype(np.int32)
            assert np.array_equal(D.col_sums_in_rows(noisy, negate), got)


def test_masked_sums_skip_entries_instead_of_multiplying_by_zero():
    """An Inf or NaN outside the selection must not reach the sum (Inf * 0 = NaN): the cursor never visits it."""
    i = np.array([0, 1, 2, 0, 2, 1], np.int32)
    p = np.array([0, 3, 5, 6], np.int32)
    x = np.array([1.0, np.inf, 2.0, np.nan, 4.0, -np.inf])
    with DeviceMatrix.from_host(i, p, x, 3, 3) as D:
        assert D.col_sums_in_rows(np.array([0, 2], np.int32)).tolist()[0] == 3.0
        got = D.col_sums_in_rows(np.array([2], np.int32))
        assert got.tolist() == [2.0, 4.0, 0.0]
        got = D.col_sums_in_rows(np.array([2], np.int32), negate=True)
        assert got[0] == np.inf and np.isnan(got[1]) and got[2] == -np.inf


def test_dense_extraction_matches_numpy():
    spec = synth.powerlaw_spec(400, 300, 30.0, 8, row_levels=3)
    i, p, x = synth.generate_host(spec)
    dense = np.zeros((spec.nrow, spec.ncol))
    dense[i, np.repeat(np.arange(spec.ncol), np.diff(p))] = x
    rng = np.random.default_rng(5)
    rows = rng.integers(0, spec.nrow, 37).astype(np.int32)  # any order, duplicates allowed (reference :85-92)
    cols = rng.integers(0, spec.ncol, 23).astype(np.int32)
    with DeviceMatrix.from_host(i, p, x, spec.nrow, spec.ncol) as D:
        assert np.array_equal(D.gather_block(rows, cols), dense[np.ix_(rows, cols)])
        assert np.array_equal(D.gather_block(None, cols), dense[:, cols])          # col(IntegerVector), :100-107
        assert np.array_equal(D.gather_block(rows, None), dense[rows, :])          # row(IntegerVector), :117-128
        assert np.array_equal(D.gather_block(None, cols[:1])[:, 0], dense[:, cols[0]])  # col(int), :95-99
        assert D.gather_block(np.zeros(0, np.int32), cols).shape == (0, 23)
