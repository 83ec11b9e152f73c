// Stand-in for <RcppArmadillo.h>/<Rcpp.h> so that the reference's hot-path sources
// (/root/reference/src/*.cpp) compile UNMODIFIED without R.  Test infrastructure only:
// it is used by oracle/Makefile to build oracle/_ref/libexref.so, the live reference the
// CPU restatement and the CUDA sampler are checked against.  It provides just the
// subset of the Rcpp / Armadillo surface those ten translation units touch.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

// ------------------------------------------------------------------ R C API subset
enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, S4SXP = 25 };

struct RNode;
typedef std::shared_ptr<RNode> SEXP;

struct RNode {
  int type = NILSXP;
  std::string klass;                 // S4 class name
  std::vector<int> ints;             // INTSXP / LGLSXP payload
  std::vector<double> reals;         // REALSXP payload
  std::vector<std::string> strs;     // STRSXP payload
  std::vector<SEXP> items;           // VECSXP payload
  std::vector<std::string> names;    // names of a VECSXP
  std::map<std::string, SEXP> attrs; // S4 slots and plain attributes
  int nrow = 0, ncol = 0;
};

inline int TYPEOF(const SEXP &x) { return x ? x->type : NILSXP; }
#define NA_INTEGER INT_MIN
#define R_FINITE(x) std::isfinite(x)
static const double R_NegInf = -std::numeric_limits<double>::infinity();
static const double R_PosInf = std::numeric_limits<double>::infinity();
static const SEXP R_NilValue = SEXP();

namespace R {
double unif_rand();
double rgamma(double shape, double scale);
double rbeta(double a, double b);
double rnbinom(double size, double prob);
}

// ------------------------------------------------------------------ Armadillo subset
namespace arma {
template <class T> struct Col {
  std::vector<T> d;
  size_t n_rows = 0, n_cols = 1, n_elem = 0;
  Col() {}
  explicit Col(size_t n) : d(n), n_rows(n), n_elem(n) {}
  Col(const std::vector<T> &v) : d(v), n_rows(v.size()), n_elem(v.size()) {}
  T &operator()(size_t i) { return d[i]; }
  const T &operator()(size_t i) const { return d[i]; }
  T &operator[](size_t i) { return d[i]; }
  const T &operator[](size_t i) const { return d[i]; }
  size_t size() const { return d.size(); }
  typename std::vector<T>::iterator begin() { return d.begin(); }
  typename std::vector<T>::iterator end() { return d.end(); }
  typename std::vector<T>::const_iterator begin() const { return d.begin(); }
  typename std::vector<T>::const_iterator end() const { return d.end(); }
  typename std::vector<T>::const_iterator cbegin() const { return d.begin(); }
  typename std::vector<T>::const_iterator cend() const { return d.end(); }
  Col &operator-=(T s) { for (auto &x : d) x -= s; return *this; }
  Col operator+(T s) const { Col r(*this); for (auto &x : r.d) x += s; return r; }
};
template <class T> struct Mat {
  std::vector<T> d;
  size_t n_rows = 0, n_cols = 0;
  Mat() {}
  Mat(size_t r, size_t c) : d(r * c), n_rows(r), n_cols(c) {}
  T &operator()(size_t i, size_t j) { return d[i + j * n_rows]; }
  const T &operator()(size_t i, size_t j) const { return d[i + j * n_rows]; }
  struct ColRef {
    Mat *m; size_t j;
    ColRef &operator=(const Col<T> &v) {
      for (size_t i = 0; i < m->n_rows; i++) (*m)(i, j) = v[i];
      return *this;
    }
  };
  ColRef col(size_t j) { return ColRef{this, j}; }
  Col<T> col(size_t j) const {
    return Col<T>(std::vector<T>(d.begin() + j * n_rows, d.begin() + (j + 1) * n_rows));
  }
  typename std::vector<T>::iterator begin_col(size_t j) { return d.begin() + j * n_rows; }
  typename std::vector<T>::iterator end_col(size_t j) { return d.begin() + (j + 1) * n_rows; }
  typename std::vector<T>::const_iterator begin_col(size_t j) const { return d.begin() + j * n_rows; }
  typename std::vector<T>::const_iterator end_col(size_t j) const { return d.begin() + (j + 1) * n_rows; }
  Mat &operator-=(T s) { for (auto &x : d) x -= s; return *this; }
};
typedef Col<int> ivec;
typedef Col<double> vec;
typedef Mat<int> imat;
typedef Mat<double> mat;
template <class M> struct zeros_impl;
template <class T> struct zeros_impl<Col<T>> { static Col<T> make(size_t n, size_t) { return Col<T>(n); } };
template <class T> struct zeros_impl<Mat<T>> { static Mat<T> make(size_t r, size_t c) { return Mat<T>(r, c); } };
template <class M> M zeros(size_t r, size_t c = 1) { return zeros_impl<M>::make(r, c); }
} // namespace arma

// ------------------------------------------------------------------ Rcpp subset
namespace Rcpp {
inline void stop(const std::string &msg) { throw std::runtime_error(msg); }
inline void warning(const std::string &msg) { std::cerr << "warning: " << msg << std::endl; }
static std::ostream &Rcout = std::cout;

inline SEXP make_node(int type) { SEXP n = std::make_shared<RNode>(); n->type = type; return n; }
inline SEXP make_real(const std::vector<double> &v) { SEXP n = make_node(REALSXP); n->reals = v; return n; }
inline SEXP make_int(const std::vector<int> &v) { SEXP n = make_node(INTSXP); n->ints = v; return n; }
inline SEXP make_lgl(bool b) { SEXP n = make_node(LGLSXP); n->ints = {b ? 1 : 0}; return n; }
inline SEXP make_str(const std::vector<std::string> &v) { SEXP n = make_node(STRSXP); n->strs = v; return n; }

inline double node_as_double(const SEXP &n) {
  if (TYPEOF(n) == REALSXP) return n->reals.at(0);
  if (TYPEOF(n) == INTSXP || TYPEOF(n) == LGLSXP) return n->ints.at(0);
  throw std::runtime_error("shim: not a number");
}

class RObject {
public:
  SEXP node;
  RObject() {}
  RObject(const SEXP &n) : node(n) {}
  operator SEXP() const { return node; }
};

class AttrProxy; // assignable view of one slot / attribute

template <class T, int TYPE> class Vector : public RObject {
  std::vector<T> &payload() const;
public:
  Vector() : RObject(make_node(TYPE)) {}
  Vector(const SEXP &n) : RObject(n) {
    if (TYPE == REALSXP && TYPEOF(n) == INTSXP) { // integer -> numeric coercion
      node = make_node(REALSXP);
      node->reals.assign(n->ints.begin(), n->ints.end());
      node->attrs = n->attrs;
    }
  }
  Vector(const RObject &o) : Vector(o.node) {}
  explicit Vector(int n) : RObject(make_node(TYPE)) { payload().resize(n); }
  template <class It> Vector(It b, It e) : RObject(make_node(TYPE)) { payload().assign(b, e); }
  T &operator[](size_t i) { return payload()[i]; }
  const T &operator[](size_t i) const { return payload()[i]; }
  int size() const { return (int)payload().size(); }
  int length() const { return size(); }
  typename std::vector<T>::iterator begin() { return payload().begin(); }
  typename std::vector<T>::iterator end() { return payload().end(); }
  typename std::vector<T>::const_iterator begin() const { return payload().begin(); }
  typename std::vector<T>::const_iterator end() const { return payload().end(); }
  AttrProxy attr(const std::string &name);
  Vector<std::string, STRSXP> names() const;
};
template <> inline std::vector<int> &Vector<int, INTSXP>::payload() const { return node->ints; }
template <> inline std::vector<double> &Vector<double, REALSXP>::payload() const { return node->reals; }
template <> inline std::vector<std::string> &Vector<std::string, STRSXP>::payload() const { return node->strs; }
typedef Vector<int, INTSXP> IntegerVector;
typedef Vector<double, REALSXP> NumericVector;
typedef Vector<std::string, STRSXP> CharacterVector;
template <class T, int TYPE> Vector<std::string, STRSXP> Vector<T, TYPE>::names() const {
  auto it = node->attrs.find("names");
  return it == node->attrs.end() ? CharacterVector() : CharacterVector(it->second);
}

class IntegerMatrix : public IntegerVector {
public:
  IntegerMatrix() {}
  IntegerMatrix(const SEXP &n) : IntegerVector(n) {}
  IntegerMatrix(const RObject &o) : IntegerVector(o.node) {}
};
class NumericMatrix : public NumericVector {
public:
  NumericMatrix() {}
  NumericMatrix(const SEXP &n) : NumericVector(n) {}
  NumericMatrix(const RObject &o) : NumericVector(o.node) {}
};

class S4;
class List;

// value read out of (or written into) a slot, an attribute or a list element
class AttrProxy {
  SEXP owner; std::string name; SEXP direct; // direct: list element (read-only)
public:
  AttrProxy(const SEXP &o, const std::string &n) : owner(o), name(n) {}
  explicit AttrProxy(const SEXP &elem) : direct(elem) {}
  SEXP get() const {
    if (direct || !owner) return direct;
    auto it = owner->attrs.find(name);
    if (it == owner->attrs.end()) throw std::runtime_error("shim: no slot/attribute '" + name + "'");
    return it->second;
  }
  operator SEXP() const { return get(); }
  operator double() const { return node_as_double(get()); }
  operator int() const { return (int)node_as_double(get()); }
  operator bool() const { return node_as_double(get()) != 0; }
  operator IntegerVector() const { return IntegerVector(get()); }
  operator NumericVector() const { return NumericVector(get()); }
  operator CharacterVector() const { return CharacterVector(get()); }
  operator IntegerMatrix() const { return IntegerMatrix(get()); }
  operator NumericMatrix() const { return NumericMatrix(get()); }
  operator S4() const;
  operator List() const;
  operator arma::vec() const { return arma::vec(NumericVector(get()).node->reals); }
  operator arma::ivec() const { return arma::ivec(get()->ints); }
  AttrProxy &operator=(const SEXP &v) { owner->attrs[name] = v; return *this; }
  AttrProxy &operator=(const RObject &v) { owner->attrs[name] = v.node; return *this; }
  AttrProxy &operator=(double v) { owner->attrs[name] = make_real({v}); return *this; }
  AttrProxy &operator=(int v) { owner->attrs[name] = make_int({v}); return *this; }
};
template <class T, int TYPE> AttrProxy Vector<T, TYPE>::attr(const std::string &name) {
  return AttrProxy(node, name);
}

class S4 : public RObject {
public:
  S4() {}
  S4(const SEXP &n) : RObject(n) {}
  S4(const RObject &o) : RObject(o.node) {}
  explicit S4(const std::string &klass) : RObject(make_node(S4SXP)) { node->klass = klass; }
  S4(const char *klass) : S4(std::string(klass)) {}
  bool is(const std::string &klass) const { return node && node->klass == klass; }
  AttrProxy slot(const std::string &name) const { return AttrProxy(node, name); }
  AttrProxy attr(const std::string &name) const { return AttrProxy(node, name); }
};

class List : public RObject {
public:
  List() : RObject(make_node(VECSXP)) {}
  List(const SEXP &n) : RObject(n) {}
  List(const RObject &o) : RObject(o.node) {}
  int size() const { return node ? (int)node->items.size() : 0; }
  int length() const { return size(); }
  AttrProxy operator[](int i) const { return AttrProxy(node->items.at(i)); }
  AttrProxy operator[](const std::string &key) const {
    for (size_t i = 0; i < node->names.size(); i++)
      if (node->names[i] == key) return AttrProxy(node->items[i]);
    throw std::runtime_error("shim: no list element '" + key + "'");
  }
  AttrProxy operator[](const char *key) const { return (*this)[std::string(key)]; }
  struct iterator {
    const List *l; int i;
    AttrProxy operator*() const { return (*l)[i]; }
    iterator &operator++() { ++i; return *this; }
    bool operator!=(const iterator &o) const { return i != o.i; }
  };
  iterator begin() const { return iterator{this, 0}; }
  iterator end() const { return iterator{this, size()}; }
  void push_back(const SEXP &x, const std::string &name = "") {
    node->items.push_back(x); node->names.push_back(name);
  }
};
inline AttrProxy::operator S4() const { return S4(get()); }
inline AttrProxy::operator List() const { return List(get()); }

template <class T> T as(const SEXP &x) { return T(x); }
template <class T> T as(const RObject &x) { return T(x.node); }

inline RObject wrap(const std::vector<int> &v) { return RObject(make_int(v)); }
inline RObject wrap(const arma::ivec &v) { return RObject(make_int(v.d)); }
inline RObject wrap(const arma::vec &v) { return RObject(make_real(v.d)); }
inline RObject wrap(const arma::imat &m) {
  SEXP n = make_int(m.d); n->nrow = (int)m.n_rows; n->ncol = (int)m.n_cols; return RObject(n);
}
inline RObject wrap(const arma::mat &m) {
  SEXP n = make_real(m.d); n->nrow = (int)m.n_rows; n->ncol = (int)m.n_cols; return RObject(n);
}
inline NumericVector rep(double x, int n) { NumericVector v(n); for (auto &e : v) e = x; return v; }
} // namespace Rcpp
