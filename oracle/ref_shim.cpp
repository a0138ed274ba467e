// oracle/_ref/liboracle_ref.so — the REFERENCE's own code as the parity oracle.
// TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
// CPU-baseline / --impl reference legs may load this library.
//
// What is compiled here (sources stay where they lie, nothing is copied):
//   $(REF)/src/example.cpp            columnSums()            example.cpp:26-32
//   $(REF)/inst/include/RcppSparse.h  Matrix::{colSums,rowSums,colMeans,rowMeans,crossprod,
//                                     transpose,InnerIterator} RcppSparse.h:131-194,218-233,375-385
// against the Rcpp stand-in in oracle/stub/ (R and Rcpp are not installed here).
//
// Two things the reference tree does NOT contain, restated below and labelled:
//   (1) R's Matrix::t for dgCMatrix, which RcppSparse.h:381-383 calls through the
//       R interpreter (Matrix is an unpinned, un-vendored dependency, DESCRIPTION:10).
//       Restated as a serial counting sort; its contract (canonical CSC of A^T:
//       Dim swapped, row indices ascending within each column, values permuted)
//       is cross-checked against scipy in tests/test_oracle.py.  PARITY UNPINNED
//       by the reference (it has no tests at all).
//   (2) A*v and A^T*v: the reference has no SpMV function.  They
//       are written with the reference's own InnerIterator in the idiom of
//       example.cpp:28-30.  PARITY UNPINNED by the reference; cross-checked vs scipy.
#include <example.cpp>  // brings in ../inst/include/RcppSparse.h exactly once (it has no include guard)

#include <cstdint>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_last_error;

RcppSparse::Matrix borrow(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz,
                          Rcpp::IntegerVector& dim_storage) {
  dim_storage = Rcpp::IntegerVector({nrow, ncol});
  return RcppSparse::Matrix(Rcpp::NumericVector::view(const_cast<double*>(x), long(nnz)),
                            Rcpp::IntegerVector::view(const_cast<int*>(i), long(nnz)),
                            Rcpp::IntegerVector::view(const_cast<int*>(p), long(ncol) + 1), dim_storage);
}

void copy_out(const Rcpp::NumericVector& v, double* out) {
  if (v.size() > 0) std::memcpy(out, v.begin(), sizeof(double) * size_t(v.size()));
}

// (1) stand-in for R's Matrix::t(dgCMatrix): serial counting sort over row indices.
Rcpp::S4 matrix_t_restatement(const Rcpp::S4& a) {
  Rcpp::NumericVector ax = a.slot("x");
  Rcpp::IntegerVector ai = a.slot("i"), ap = a.slot("p"), adim = a.slot("Dim");
  const int nrow = adim[0], ncol = adim[1];
  const long nnz = ax.size();
  Rcpp::IntegerVector tp(long(nrow) + 1), ti(nnz), tdim({ncol, nrow});
  Rcpp::NumericVector tx(nnz);
  for (long k = 0; k < nnz; ++k) tp[ai[k] + 1] += 1;          // entries per row
  for (int r = 0; r < nrow; ++r) tp[r + 1] += tp[r];           // exclusive scan -> new column pointers
  std::vector<int> fill(tp.begin(), tp.begin() + nrow);        // next free slot of each new column
  for (int c = 0; c < ncol; ++c)                               // source columns ascending => stable
    for (int k = ap[c]; k < ap[c + 1]; ++k) {
      const int slot = fill[ai[k]]++;
      ti[slot] = c;
      tx[slot] = ax[k];
    }
  Rcpp::S4 t(std::string("dgCMatrix"));
  t.slot("i") = ti;
  t.slot("p") = tp;
  t.slot("x") = tx;
  t.slot("Dim") = tdim;
  return t;
}

struct RegisterMatrixT {
  RegisterMatrixT() { Rcpp::stub_register_function("Matrix", "t", matrix_t_restatement); }
} g_register_matrix_t;

template <typename Body>
int guarded(Body body) {
  try {
    body();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  } catch (...) {
    g_last_error = "unknown C++ exception";
    return -1;
  }
}

}  // namespace

extern "C" {

const char* oref_last_error(void) { return g_last_error.c_str(); }

// reference src/example.cpp:26-32, run verbatim
int oref_columnSums(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, double* out) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    RcppSparse::Matrix A = borrow(i, p, x, nrow, ncol, nnz, d);
    copy_out(columnSums(A), out);
  });
}

// reference RcppSparse.h:131-137
int oref_colSums(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, double* out) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    copy_out(borrow(i, p, x, nrow, ncol, nnz, d).colSums(), out);
  });
}

// reference RcppSparse.h:138-144
int oref_rowSums(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, double* out) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    copy_out(borrow(i, p, x, nrow, ncol, nnz, d).rowSums(), out);
  });
}

// reference RcppSparse.h:158-194, run verbatim (dense n x n result, column-major; the OpenMP pragma is inert:
// this library is built without -fopenmp).  The reference's inner do-while reads i[] one element past a column's
// end before testing the bound (:178-180, :182-184); callers pass an i array with one spare element.
int oref_crossprod(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, double* out) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    Rcpp::NumericMatrix r = borrow(i, p, x, nrow, ncol, nnz, d).crossprod();
    if (ncol > 0) std::memcpy(out, r.begin(), sizeof(double) * size_t(ncol) * size_t(ncol));
  });
}

// reference RcppSparse.h:145-150
int oref_colMeans(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, double* out) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    copy_out(borrow(i, p, x, nrow, ncol, nnz, d).colMeans(), out);
  });
}

// reference RcppSparse.h:151-156
int oref_rowMeans(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, double* out) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    copy_out(borrow(i, p, x, nrow, ncol, nnz, d).rowMeans(), out);
  });
}

// reference RcppSparse.h:375-385 run verbatim; the "Matrix::t" it calls is restatement (1) above.
int oref_transpose(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, int* p_out,
                   int* i_out, double* x_out, int* dim_out) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    RcppSparse::Matrix T = borrow(i, p, x, nrow, ncol, nnz, d).transpose();
    std::memcpy(p_out, T.p.begin(), sizeof(int) * size_t(T.p.size()));
    if (nnz > 0) {
      std::memcpy(i_out, T.i.begin(), sizeof(int) * size_t(nnz));
      std::memcpy(x_out, T.x.begin(), sizeof(double) * size_t(nnz));
    }
    dim_out[0] = T.Dim[0];
    dim_out[1] = T.Dim[1];
  });
}

// (2) y = A v : restatement in the reference's iterator idiom (example.cpp:28-30 with the
//     scatter shape of RcppSparse.h:140-142).  y has nrow entries, v has ncol.
int oref_spmv(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, const double* v,
              double* y) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    RcppSparse::Matrix A = borrow(i, p, x, nrow, ncol, nnz, d);
    Rcpp::NumericVector acc(A.rows());
    for (size_t col = 0; col < A.cols(); ++col)
      for (RcppSparse::Matrix::InnerIterator it(A, col); it; ++it) acc(it.row()) += it.value() * v[col];
    copy_out(acc, y);
  });
}

// (2) y = A^T v : same idiom with the gather shape of RcppSparse.h:133-135.  y has ncol entries, v has nrow.
int oref_spmv_t(const int* i, const int* p, const double* x, int nrow, int ncol, int64_t nnz, const double* v,
                double* y) {
  return guarded([&] {
    Rcpp::IntegerVector d;
    RcppSparse::Matrix A = borrow(i, p, x, nrow, ncol, nnz, d);
    Rcpp::NumericVector acc(A.cols());
    for (size_t col = 0; col < A.cols(); ++col)
      for (RcppSparse::Matrix::InnerIterator it(A, col); it; ++it) acc(col) += it.value() * v[it.row()];
    copy_out(acc, y);
  });
}

}  // extern "C"
