// Rcpp stand-in, part 2 of 2 — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Just enough of the Rcpp surface for /root/reference/inst/include/RcppSparse.h
// and /root/reference/src/example.cpp to compile UNMODIFIED with plain g++.
// Containers only; no arithmetic on matrix data happens in this file.
//
// Semantics kept from real Rcpp because the reference depends on them:
//  * vectors are HANDLES: copying a NumericVector aliases the same storage
//    (reference vignettes/Documentation.Rmd:325-347, "Reference vs. Copy");
//  * Vector(n) zero-fills (reference RcppSparse.h:132,139 accumulate into it);
//  * operator[] is unchecked, operator() is bounds-checked and throws
//    Rcpp::index_out_of_bounds (reference RcppSparse.h:135,142 use "sums(col)");
//  * Vector::view() aliases caller memory without copying — what Rcpp does with
//    the SEXP slots of a dgCMatrix (reference RcppSparse.h:34-41, README.md:7-9).
//  * S4 is a bag of named slots; Environment/Function exist so that
//    Matrix::transpose() (reference RcppSparse.h:375-385) can call "Matrix::t".
//    The callee is whatever the test harness registered under that name — R's
//    Matrix package is not part of the reference tree, see oracle/ref_shim.cpp.
#ifndef ORACLE_STUB_RCPP_H
#define ORACLE_STUB_RCPP_H

#include "RcppCommon.h"

#include <algorithm>
#include <functional>
#include <initializer_list>
#include <iterator>
#include <map>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace Rcpp {

class index_out_of_bounds : public std::out_of_range {
public:
  explicit index_out_of_bounds(const std::string& what) : std::out_of_range(what) {}
};

template <typename T>
class Vector {
public:
  typedef T value_type;

  Vector() : base_(nullptr), len_(0) {}
  Vector(long n) : owned_(std::make_shared<std::vector<T>>(size_t(n < 0 ? 0 : n), T(0))) { rebind(); }
  Vector(long n, T fill) : owned_(std::make_shared<std::vector<T>>(size_t(n < 0 ? 0 : n), fill)) { rebind(); }
  Vector(std::initializer_list<T> init) : owned_(std::make_shared<std::vector<T>>(init)) { rebind(); }

  // Non-owning alias of memory someone else keeps alive (an R vector, a numpy array).
  static Vector view(T* data, long n) {
    Vector v;
    v.base_ = data;
    v.len_ = n;
    return v;
  }

  long size() const { return len_; }
  long length() const { return len_; }

  T& operator[](long k) { return base_[k]; }
  const T& operator[](long k) const { return base_[k]; }

  T& operator()(std::size_t k) {
    if (k >= std::size_t(len_)) throw index_out_of_bounds("Index out of bounds");
    return base_[k];
  }
  const T& operator()(std::size_t k) const {
    if (k >= std::size_t(len_)) throw index_out_of_bounds("Index out of bounds");
    return base_[k];
  }

  T* begin() { return base_; }
  T* end() { return base_ + len_; }
  const T* begin() const { return base_; }
  const T* end() const { return base_ + len_; }

private:
  void rebind() {
    base_ = owned_->data();
    len_ = long(owned_->size());
  }
  std::shared_ptr<std::vector<T>> owned_;
  T* base_;
  long len_;
};

typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;

template <typename T>
Vector<T> clone(const Vector<T>& src) {
  Vector<T> out(src.size());
  std::copy(src.begin(), src.end(), out.begin());
  return out;
}

// Column-major dense matrix; only what the (out-of-scope) dense accessors of the
// reference header need in order to parse and run.
class NumericMatrix {
public:
  class Strip {  // one column or one row, assignable from a NumericVector
  public:
    Strip(double* first, long count, long step) : first_(first), count_(count), step_(step) {}
    Strip& operator=(const NumericVector& v) {
      for (long k = 0; k < count_ && k < v.size(); ++k) first_[k * step_] = v[k];
      return *this;
    }
  private:
    double* first_;
    long count_, step_;
  };

  NumericMatrix() : nr_(0), nc_(0) {}
  NumericMatrix(long nr, long nc) : cells_(nr * nc), nr_(nr), nc_(nc) {}
  double& operator()(long r, long c) { return cells_[c * nr_ + r]; }
  const double& operator()(long r, long c) const { return cells_[c * nr_ + r]; }
  Strip column(long c) { return Strip(cells_.begin() + c * nr_, nr_, 1); }
  Strip row(long r) { return Strip(cells_.begin() + r, nc_, nr_); }
  long nrow() const { return nr_; }
  long ncol() const { return nc_; }
  double* begin() { return cells_.begin(); }

private:
  NumericVector cells_;
  long nr_, nc_;
};

namespace stubdetail {
struct SexpRec {
  std::string klass;
  std::map<std::string, NumericVector> dbl;
  std::map<std::string, IntegerVector> itg;
};
}  // namespace stubdetail

class S4 {
public:
  class Slot {
  public:
    Slot(stubdetail::SexpRec* rec, const std::string& name) : rec_(rec), name_(name) {}
    operator NumericVector() const {
      std::map<std::string, NumericVector>::const_iterator it = rec_->dbl.find(name_);
      if (it == rec_->dbl.end()) throw std::invalid_argument("no numeric slot " + name_);
      return it->second;
    }
    operator IntegerVector() const {
      std::map<std::string, IntegerVector>::const_iterator it = rec_->itg.find(name_);
      if (it == rec_->itg.end()) throw std::invalid_argument("no integer slot " + name_);
      return it->second;
    }
    Slot& operator=(const NumericVector& v) {
      rec_->dbl[name_] = v;
      return *this;
    }
    Slot& operator=(const IntegerVector& v) {
      rec_->itg[name_] = v;
      return *this;
    }
  private:
    stubdetail::SexpRec* rec_;
    std::string name_;
  };

  S4() : rec_(std::make_shared<stubdetail::SexpRec>()) {}
  explicit S4(const std::string& klass) : rec_(std::make_shared<stubdetail::SexpRec>()) { rec_->klass = klass; }
  // From a raw SEXP: alias, do not own (R would own it).
  S4(SEXP raw) : rec_(raw, [](stubdetail::SexpRec*) {}) {}

  bool hasSlot(const std::string& name) const {
    return rec_->dbl.count(name) != 0 || rec_->itg.count(name) != 0;
  }
  Slot slot(const std::string& name) const { return Slot(rec_.get(), name); }
  SEXP get() const { return rec_.get(); }
  const std::string& klass() const { return rec_->klass; }

private:
  std::shared_ptr<stubdetail::SexpRec> rec_;
};

// Rcpp::_["name"] = value
struct NamedS4 {
  std::string name;
  S4 value;
};
struct ArgName {
  std::string name;
  NamedS4 operator=(const S4& v) const { return NamedS4{name, v}; }
};
struct NamedPlaceHolder {
  ArgName operator[](const std::string& name) const { return ArgName{name}; }
};
static const NamedPlaceHolder _ = NamedPlaceHolder();

typedef std::function<S4(const S4&)> UnaryS4Fn;

namespace stubdetail {
inline std::map<std::string, UnaryS4Fn>& callee_table() {
  static std::map<std::string, UnaryS4Fn> table;
  return table;
}
}  // namespace stubdetail

// Test harness hook: make `Environment::namespace_env(ns)[name]` callable.
inline void stub_register_function(const std::string& ns, const std::string& name, UnaryS4Fn fn) {
  stubdetail::callee_table()[ns + "::" + name] = fn;
}

class Function {
public:
  Function() {}
  explicit Function(const std::string& key) : key_(key) {}
  S4 operator()(const NamedS4& arg) const {
    std::map<std::string, UnaryS4Fn>::const_iterator it = stubdetail::callee_table().find(key_);
    if (it == stubdetail::callee_table().end())
      throw std::runtime_error("Rcpp stand-in: no function registered as " + key_);
    return it->second(arg.value);
  }
private:
  std::string key_;
};

class Environment {
public:
  static Environment namespace_env(const std::string& ns) {
    Environment e;
    e.ns_ = ns;
    return e;
  }
  Function operator[](const std::string& name) const { return Function(ns_ + "::" + name); }
private:
  std::string ns_;
};

}  // namespace Rcpp

#endif
