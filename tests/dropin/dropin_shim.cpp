// Test harness for the drop-in C++ header: user-level code exactly as a downstream Rcpp package
// would write it against RcppSparse::Matrix, compiled against the Rcpp stand-in (R and Rcpp are not
// installed here) and linked to libsparse_b200.  The extern "C" wrappers only move raw arrays in
// and out so that pytest can drive it with ctypes.
#include <RcppSparse.h>  // the drop-in header from include/

#include <cstdint>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

// The package's exported function (reference src/example.cpp:26-32) in the drop-in build: the sweep is
// one call on the class instead of a serial InnerIterator loop.
Rcpp::NumericVector columnSums(RcppSparse::Matrix& A) { return A.colSums(); }

// A downstream user's own iterator loop keeps working unchanged (host side, reference idiom).
static Rcpp::NumericVector column_sums_by_iterator(RcppSparse::Matrix& A) {
  Rcpp::NumericVector sums(A.cols());
  for (size_t col = 0; col < A.cols(); ++col)
    for (RcppSparse::Matrix::InnerIterator it(A, col); it; ++it) sums(col) += it.value();
  return sums;
}

namespace {
thread_local std::string g_err;

RcppSparse::Matrix view(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz) {
  return RcppSparse::Matrix(Rcpp::NumericVector::view(const_cast<double*>(x), long(nnz)),
                            Rcpp::IntegerVector::view(const_cast<int*>(i), long(nnz)),
                            Rcpp::IntegerVector::view(const_cast<int*>(p), long(ncol) + 1),
                            Rcpp::IntegerVector({nrow, ncol}));
}
void out(const Rcpp::NumericVector& v, double* dst) {
  if (v.size() > 0) std::memcpy(dst, v.begin(), sizeof(double) * size_t(v.size()));
}
template <typename F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::invalid_argument& e) {
    g_err = std::string("invalid_argument: ") + e.what();
    return 1;
  } catch (const std::exception& e) {
    g_err = std::string("runtime_error: ") + e.what();
    return 2;
  }
}
}  // namespace

extern "C" {
const char* dropin_last_error() { return g_err.c_str(); }

// op: 0 columnSums (exported fn), 1 colSums, 2 rowSums, 3 colMeans, 4 rowMeans, 5 iterator loop on the host
int dropin_reduce(int op, const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, double* dst) {
  return guarded([&] {
    RcppSparse::Matrix A = view(i, p, x, nrow, ncol, nnz);
    switch (op) {
      case 0: out(columnSums(A), dst); break;
      case 1: out(A.colSums(), dst); break;
      case 2: out(A.rowSums(), dst); break;
      case 3: out(A.colMeans(), dst); break;
      case 4: out(A.rowMeans(), dst); break;
      default: out(column_sums_by_iterator(A), dst); break;
    }
  });
}
int dropin_crossprod(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, double* dst) {
  return guarded([&] {
    RcppSparse::Matrix A = view(i, p, x, nrow, ncol, nnz);
    Rcpp::NumericMatrix r = A.crossprod();
    if (r.nrow() != ncol || r.ncol() != ncol) throw std::runtime_error("crossprod: wrong shape");
    if (ncol > 0) std::memcpy(dst, r.begin(), sizeof(double) * size_t(ncol) * size_t(ncol));
  });
}
int dropin_spmv(int transposed, const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz,
                const double* v, double* y) {
  return guarded([&] {
    RcppSparse::Matrix A = view(i, p, x, nrow, ncol, nnz);
    Rcpp::NumericVector vv = Rcpp::NumericVector::view(const_cast<double*>(v), transposed ? nrow : ncol);
    out(transposed ? A.spmv_t(vv) : A.spmv(vv), y);
  });
}
int dropin_transpose(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, int* tp, int* ti,
                     double* tx, int* tdim) {
  return guarded([&] {
    RcppSparse::Matrix A = view(i, p, x, nrow, ncol, nnz);
    RcppSparse::Matrix T = A.t();
    std::memcpy(tp, T.p.begin(), sizeof(int) * size_t(T.p.size()));
    if (nnz > 0) {
      std::memcpy(ti, T.i.begin(), sizeof(int) * size_t(nnz));
      std::memcpy(tx, T.x.begin(), sizeof(double) * size_t(nnz));
    }
    tdim[0] = T.Dim[0];
    tdim[1] = T.Dim[1];
    // round trip through wrap() and the Exporter, as Rcpp::as<RcppSparse::Matrix> would do
    Rcpp::S4 s = T.wrap();
    RcppSparse::Matrix back = Rcpp::traits::Exporter<RcppSparse::Matrix>(s.get()).get();
    if (back.x.size() != T.x.size() || back.rows() != T.rows()) throw std::runtime_error("wrap/as round trip");
  });
}
// The reference reads x live on every call: an in-place edit is seen by the next call without any refresh.  A
// resident Matrix keeps its mirror (copies share it; clone() does not) and needs refresh() after an in-place edit.
int dropin_alias_semantics(const int* i, const int* p, double* x, int nrow, int ncol, int64_t nnz, double* before,
                           double* after, double* stale, double* refreshed) {
  return guarded([&] {
    RcppSparse::Matrix A = view(i, p, x, nrow, ncol, nnz);
    RcppSparse::Matrix B = A;  // aliases the same vectors and the same device state
    out(B.colSums(), before);
    x[0] += 1000.0;
    out(B.colSums(), after);  // live read, like the reference's loop over the R vector
    A.resident();
    if (!B.is_resident()) throw std::runtime_error("copies share the device state");
    out(A.colSums(), after);
    x[0] += 1000.0;
    out(B.colSums(), stale);  // resident: the mirror still holds the old value
    A.refresh();
    out(B.colSums(), refreshed);
  });
}
// public members may be re-pointed after construction (vignette: "m2.x = x; m2.i = i; ..."): the mirror follows
int dropin_repointed_members(const int* i, const int* p, const double* x1, const double* x2, int nrow, int ncol,
                             int64_t nnz, double* sums1, double* sums2) {
  return guarded([&] {
    RcppSparse::Matrix A;  // default-constructed, then filled slot by slot
    A.i = Rcpp::IntegerVector::view(const_cast<int*>(i), long(nnz));
    A.p = Rcpp::IntegerVector::view(const_cast<int*>(p), long(ncol) + 1);
    A.Dim = Rcpp::IntegerVector({nrow, ncol});
    A.x = Rcpp::NumericVector::view(const_cast<double*>(x1), long(nnz));
    out(A.colSums(), sums1);
    A.x = Rcpp::NumericVector::view(const_cast<double*>(x2), long(nnz));  // another vector: new upload
    out(A.colSums(), sums2);
  });
}
// ---- host-side accessors and cursors (no device involved) ---------------------------------------------------
int dropin_dense_parts(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, int r, int c,
                       const int* rows, int nrows, const int* cols, int ncols, double* row_out, double* col_out,
                       double* block_out, double* rowsel_out, double* colsel_out, double* cols_out, double* rows_out) {
  return guarded([&] {
    RcppSparse::Matrix A = view(i, p, x, nrow, ncol, nnz);
    Rcpp::IntegerVector rv = Rcpp::IntegerVector::view(const_cast<int*>(rows), nrows);
    Rcpp::IntegerVector cv = Rcpp::IntegerVector::view(const_cast<int*>(cols), ncols);
    out(A.row(r), row_out);
    out(A.col(c), col_out);
    out(A(r, cv), rowsel_out);
    out(A(rv, c), colsel_out);
    Rcpp::NumericMatrix b = A(rv, cv);
    for (int k = 0; k < ncols; ++k)
      for (int j = 0; j < nrows; ++j) block_out[size_t(k) * nrows + j] = b(j, k);
    Rcpp::NumericMatrix cm = A.col(cv);  // nrow x ncols
    for (int k = 0; k < ncols; ++k)
      for (int j = 0; j < nrow; ++j) cols_out[size_t(k) * nrow + j] = cm(j, k);
    Rcpp::NumericMatrix rm = A.row(rv);  // nrows x ncol
    for (int k = 0; k < ncol; ++k)
      for (int j = 0; j < nrows; ++j) rows_out[size_t(k) * nrows + j] = rm(j, k);
  });
}
// entries of column `col` whose rows are in / not in the sorted list s; the rows with / without an entry
int dropin_range_cursors(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, int col,
                         const unsigned* s, int ns, int* in_rows, double* in_vals, int* n_in, int* out_rows,
                         double* out_vals, int* n_out, unsigned* inner, int* n_inner, unsigned* empty, int* n_empty) {
  return guarded([&] {
    RcppSparse::Matrix A = view(i, p, x, nrow, ncol, nnz);
    std::vector<unsigned int> sv(s, s + ns);
    int k = 0;
    for (RcppSparse::Matrix::InnerIteratorInRange it(A, col, sv); it; ++it, ++k) {
      if (it.col() != col) throw std::runtime_error("InRange: col()");
      in_rows[k] = it.row();
      in_vals[k] = it.value();
    }
    *n_in = k;
    k = 0;
    for (RcppSparse::Matrix::InnerIteratorNotInRange it(A, col, sv); it; ++it, ++k) {
      out_rows[k] = it.row();
      out_vals[k] = it.value();
    }
    *n_out = k;
    std::vector<unsigned int> a = A.InnerIndices(col), e = A.emptyInnerIndices(col);
    std::copy(a.begin(), a.end(), inner);
    std::copy(e.begin(), e.end(), empty);
    *n_inner = int(a.size());
    *n_empty = int(e.size());
  });
}
int dropin_row_cursor(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, int row, int* cols,
                      double* vals, int* n, int* symmetric) {
  return guarded([&] {
    RcppSparse::Matrix A = view(i, p, x, nrow, ncol, nnz);
    int k = 0;
    for (RcppSparse::Matrix::InnerRowIterator it(A, row); it; ++it, ++k) {
      if (it.row() != row) throw std::runtime_error("InnerRowIterator: row()");
      cols[k] = it.col();
      vals[k] = it.value();
    }
    *n = k;
    *symmetric = A.isAppxSymmetric() ? 1 : 0;
  });
}
int dropin_missing_slot() {
  return guarded([&] {
    Rcpp::S4 s(std::string("dgCMatrix"));
    s.slot("x") = Rcpp::NumericVector(1);
    RcppSparse::Matrix bad(s);
  });
}
}
