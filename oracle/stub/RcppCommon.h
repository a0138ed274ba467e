// Rcpp stand-in, part 1 of 2 — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// R, Rcpp and libR are not installed in this image, so the reference header
// (/root/reference/inst/include/RcppSparse.h) cannot be compiled as shipped.
// This file and Rcpp.h provide just the Rcpp names that header mentions, so
// that it compiles UNMODIFIED and its own loops can serve as the parity oracle.
// Nothing here does arithmetic on matrix data: containers only.
//
// <RcppCommon.h> in real Rcpp declares the traits templates a user may
// specialise before <Rcpp.h> is included (reference RcppSparse.h:1-16 relies on
// that order). We mirror exactly that: the primary Exporter template lives here.
#ifndef ORACLE_STUB_RCPPCOMMON_H
#define ORACLE_STUB_RCPPCOMMON_H

#include <cstddef>

namespace Rcpp {
namespace stubdetail {
struct SexpRec;  // a heap record standing in for an R object
}
}  // namespace Rcpp

typedef Rcpp::stubdetail::SexpRec* SEXP;

namespace Rcpp {
namespace traits {
template <typename T>
class Exporter;  // primary template: specialised by the reference header
}  // namespace traits
}  // namespace Rcpp

#endif
