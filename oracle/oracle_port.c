/* oracle/liboracle_port.so — plain-C restatement of the reference's serial hot path.
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product (libsparse_b200)
 * never links, loads or calls it.
 *
 * Reference = /root/reference (zdebruine/RcppSparse).  Each function names the lines
 * it follows.  Loop order and accumulation order are the reference's, so FP64 results
 * are bit-identical to oracle/_ref (the reference header compiled verbatim); that
 * equality is asserted by tests/test_oracle.py wherever /root/reference is present.
 *
 * Pinning: the reference ships NO tests, golden vectors or known-answer fixtures
 * (SURVEY.md section 4).  The port is pinned instead against (a) oracle/_ref, i.e. the
 * reference's own compiled loops, and (b) the literal 5x5 matrix of
 * vignettes/Documentation.Rmd:213-216 (tests/golden/vignette_5x5.json).  For
 * transpose() and the two SpMV sweeps the reference tree holds no arithmetic at all
 * (RcppSparse.h:381-383 calls R's Matrix::t; no SpMV exists) => "parity unpinned" by
 * the reference; these are cross-checked against scipy.sparse in the tests.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* RcppSparse.h:131-137 — per-column serial left-to-right sum; never reads i. */
void oport_colSums(const int* p, const double* x, int ncol, double* sums) {
  for (int col = 0; col < ncol; ++col) {
    double acc = 0.0;
    for (int j = p[col]; j < p[col + 1]; ++j) acc += x[j];
    sums[col] = acc;
  }
}

/* src/example.cpp:26-32 — the same sums through the InnerIterator protocol
 * (index = p[col], max_index = p[col+1], test, value(), ++; RcppSparse.h:220-226). */
void oport_columnSums(const int* p, const double* x, int ncol, double* sums) {
  for (int col = 0; col < ncol; ++col) {
    int index = p[col];
    const int max_index = p[col + 1];
    sums[col] = 0.0;
    for (; index < max_index; ++index) sums[col] += x[index];
  }
}

/* RcppSparse.h:138-144 — scatter-add in storage order (per row: ascending column). */
void oport_rowSums(const int* i, const int* p, const double* x, int nrow, int ncol, double* sums) {
  memset(sums, 0, sizeof(double) * (size_t)nrow);
  for (int col = 0; col < ncol; ++col)
    for (int j = p[col]; j < p[col + 1]; ++j) sums[i[j]] += x[j];
}

/* RcppSparse.h:158-194 — dense A^T A (n x n, column-major): for every pair of columns a two-pointer merge of
 * their sorted row lists, products added in ascending row order; the lower triangle is a copy of the upper.
 * Same arithmetic as the reference, with the bound tested BEFORE the index is read (the reference reads i[]
 * one past the column end, :178-184; harmless there, undefined here). */
void oport_crossprod(const int* i, const int* p, const double* x, int ncol, double* res) {
  memset(res, 0, sizeof(double) * (size_t)ncol * (size_t)ncol);
  for (int c1 = 0; c1 < ncol; ++c1) {
    for (int c2 = c1; c2 < ncol; ++c2) {
      double acc = 0.0;
      int a = p[c1], b = p[c2];
      const int ae = p[c1 + 1], be = p[c2 + 1];
      while (a < ae && b < be) {
        if (i[a] == i[b]) {
          acc += x[a] * x[b];
          ++a;
          ++b;
        } else if (i[a] < i[b]) {
          ++a;
        } else {
          ++b;
        }
      }
      res[(size_t)c2 * ncol + c1] = acc;
      res[(size_t)c1 * ncol + c2] = acc;
    }
  }
}

/* RcppSparse.h:145-150 — colSums then true division by Dim[0] (int promoted to double). */
void oport_colMeans(const int* p, const double* x, int nrow, int ncol, double* means) {
  oport_colSums(p, x, ncol, means);
  for (int c = 0; c < ncol; ++c) means[c] = means[c] / nrow;
}

/* RcppSparse.h:151-156 — rowSums then division by Dim[1]. */
void oport_rowMeans(const int* i, const int* p, const double* x, int nrow, int ncol, double* means) {
  oport_rowSums(i, p, x, nrow, ncol, means);
  for (int r = 0; r < nrow; ++r) means[r] = means[r] / ncol;
}

/* RcppSparse.h:375-385 delegates to R's Matrix::t (not in the reference tree).  Output
 * contract restated: canonical CSC of A^T — p_out[nrow+1], i_out = source column ids
 * ascending within each new column, x_out permuted, no arithmetic on values.
 * Serial counting sort: count rows, exclusive scan, in-order (hence stable) scatter. */
int oport_transpose(const int* i, const int* p, const double* x, int nrow, int ncol, int* p_out, int* i_out,
                    double* x_out) {
  const int64_t nnz = p[ncol];
  memset(p_out, 0, sizeof(int) * ((size_t)nrow + 1));
  for (int64_t k = 0; k < nnz; ++k) p_out[i[k] + 1] += 1;
  for (int r = 0; r < nrow; ++r) p_out[r + 1] += p_out[r];
  int* next = (int*)malloc(sizeof(int) * ((size_t)nrow + 1));
  if (!next) return -1;
  memcpy(next, p_out, sizeof(int) * (size_t)nrow);
  for (int c = 0; c < ncol; ++c)
    for (int k = p[c]; k < p[c + 1]; ++k) {
      const int slot = next[i[k]]++;
      i_out[slot] = c;
      x_out[slot] = x[k];
    }
  free(next);
  return 0;
}

/* y = A v in the idiom of example.cpp:28-30 with the scatter of RcppSparse.h:140-142
 * (the reference has no SpMV function).  Separate multiply then add,
 * as a CRAN-flag x86-64 build does (no FMA contraction at baseline x86-64). */
void oport_spmv(const int* i, const int* p, const double* x, int nrow, int ncol, const double* v, double* y) {
  memset(y, 0, sizeof(double) * (size_t)nrow);
  for (int col = 0; col < ncol; ++col)
    for (int j = p[col]; j < p[col + 1]; ++j) y[i[j]] += x[j] * v[col];
}

/* y = A^T v, gather shape of RcppSparse.h:133-135. */
void oport_spmv_t(const int* i, const int* p, const double* x, int ncol, const double* v, double* y) {
  for (int col = 0; col < ncol; ++col) {
    double acc = 0.0;
    for (int j = p[col]; j < p[col + 1]; ++j) acc += x[j] * v[i[j]];
    y[col] = acc;
  }
}

/* Per-output tolerance denominators used by the parity tests (north_star: |delta| <=
 * 1e-12 * sum|a_ij| over the entries feeding each output).  Not part of the reference. */
void oport_abs_colSums(const int* p, const double* x, int ncol, double* out) {
  for (int col = 0; col < ncol; ++col) {
    double acc = 0.0;
    for (int j = p[col]; j < p[col + 1]; ++j) acc += x[j] < 0 ? -x[j] : x[j];
    out[col] = acc;
  }
}
