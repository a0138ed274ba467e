// TEST INFRASTRUCTURE ONLY: the two Rcpp names r-package/src/exports.cpp uses that the reference header does
// not (and oracle/stub therefore lacks), so that the glue can be type-checked without R.  No behaviour.
#pragma once
#include <Rcpp.h>
namespace Rcpp {
template <typename T>
class XPtr {
public:
  explicit XPtr(T* p, bool set_delete_finalizer = true) : p_(p) { (void)set_delete_finalizer; }
  XPtr(SEXP s) : p_(reinterpret_cast<T*>(s)) {}
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }
  operator SEXP() const { return reinterpret_cast<SEXP>(p_); }
private:
  T* p_;
};
}  // namespace Rcpp
