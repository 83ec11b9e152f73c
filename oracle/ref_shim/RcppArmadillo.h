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
  AttrProxy attr(const std::string &name) const;
  Vector<std::string, STRSXP> names() const;
  template <class... Args> static Vector create(Args... args) {
    Vector v;
    const T vals[] = {T(args)...};
    v.payload().assign(vals, vals + sizeof...(Args));
    return v;
  }
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

// matrices: column-major payload, dimensions in the node (dimnames live in the attributes
// "rownames" / "colnames": see rownames() / colnames() below)
template <class Base, class T> class MatrixT : public Base {
public:
  MatrixT() {}
  MatrixT(const SEXP &n) : Base(n) {}
  MatrixT(const RObject &o) : Base(o.node) {}
  MatrixT(int nrow, int ncol) : Base(nrow * ncol) { this->node->nrow = nrow; this->node->ncol = ncol; }
  int nrow() const { return this->node->nrow; }
  int ncol() const { return this->node->ncol; }
  T &operator()(int i, int j) { return (*this)[(size_t)i + (size_t)j * this->node->nrow]; }
  const T &operator()(int i, int j) const { return (*this)[(size_t)i + (size_t)j * this->node->nrow]; }
  struct RowProxy {
    MatrixT *m; int i;
    template <class V> RowProxy &operator=(const V &v) {
      for (int j = 0; j < m->ncol(); j++) (*m)(i, j) = (T)v[j];
      return *this;
    }
  };
  RowProxy row(int i) { return RowProxy{this, i}; }
};
class IntegerMatrix : public MatrixT<IntegerVector, int> {
public:
  using MatrixT<IntegerVector, int>::MatrixT;
  IntegerMatrix() {}
  IntegerMatrix operator-(int s) const { // NA stays NA, as in R
    IntegerMatrix r(nrow(), ncol());
    for (int k = 0; k < size(); k++) r[k] = (*this)[k] == NA_INTEGER ? NA_INTEGER : (*this)[k] - s;
    r.node->attrs = node->attrs;
    return r;
  }
};
class NumericMatrix : public MatrixT<NumericVector, double> {
public:
  using MatrixT<NumericVector, double>::MatrixT;
  NumericMatrix() {}
};
inline IntegerVector operator-(const IntegerVector &v, int s) {
  IntegerVector r(v.size());
  for (int k = 0; k < v.size(); k++) r[k] = v[k] == NA_INTEGER ? NA_INTEGER : v[k] - s;
  return r;
}

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
  operator arma::imat() const {
    SEXP n = get(); arma::imat m(n->nrow, n->ncol); m.d = n->ints; return m;
  }
  operator arma::mat() const {
    SEXP n = NumericVector(get()).node; arma::mat m(get()->nrow, get()->ncol); m.d = n->reals; return m;
  }
  operator std::string() const { return get()->strs.at(0); }
  AttrProxy &operator=(const AttrProxy &v) { owner->attrs[name] = v.get(); return *this; } // slot <- slot
  AttrProxy &operator=(const SEXP &v) { // (an attribute set to NULL is removed, as in R)
    if (v || owner->type == S4SXP) owner->attrs[name] = v; else owner->attrs.erase(name);
    return *this;
  }
  AttrProxy &operator=(const RObject &v) { owner->attrs[name] = v.node; return *this; }
  AttrProxy &operator=(double v) { owner->attrs[name] = make_real({v}); return *this; }
  AttrProxy &operator=(int v) { owner->attrs[name] = make_int({v}); return *this; }
  AttrProxy &operator=(bool v) { owner->attrs[name] = make_lgl(v); return *this; }
};
template <class T, int TYPE> AttrProxy Vector<T, TYPE>::attr(const std::string &name) const {
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
  // element by name: readable like an AttrProxy, assignable (appends when the name is new)
  struct NameProxy {
    SEXP list; std::string key;
    int find() const {
      for (size_t i = 0; i < list->names.size(); i++) if (list->names[i] == key) return (int)i;
      return -1;
    }
    AttrProxy read() const {
      const int i = find();
      if (i < 0) throw std::runtime_error("shim: no list element '" + key + "'");
      return AttrProxy(list->items[i]);
    }
    template <class T> operator T() const { return (T)read(); }
    NameProxy &operator=(const RObject &v) {
      const int i = find();
      if (i >= 0) list->items[i] = v.node; else { list->items.push_back(v.node); list->names.push_back(key); }
      return *this;
    }
  };
  NameProxy operator[](const std::string &key) const { return NameProxy{node, key}; }
  NameProxy operator[](const char *key) const { return NameProxy{node, std::string(key)}; }
  CharacterVector names() const { return CharacterVector(make_str(node->names)); }
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

template <class T> struct as_impl { static T from(const SEXP &x) { return T(x); } };
template <> struct as_impl<double> { static double from(const SEXP &x) { return node_as_double(x); } };
template <> struct as_impl<int> { static int from(const SEXP &x) { return (int)node_as_double(x); } };
template <> struct as_impl<bool> { static bool from(const SEXP &x) { return node_as_double(x) != 0; } };
template <> struct as_impl<std::string> { static std::string from(const SEXP &x) { return x->strs.at(0); } };
template <> struct as_impl<arma::ivec> { static arma::ivec from(const SEXP &x) { return arma::ivec(x->ints); } };
template <> struct as_impl<arma::imat> {
  static arma::imat from(const SEXP &x) { arma::imat m(x->nrow, x->ncol); m.d = x->ints; return m; }
};
template <class T> T as(const SEXP &x) { return as_impl<T>::from(x); }
template <class T> T as(const RObject &x) { return as_impl<T>::from(x.node); }
template <class T> T as(const AttrProxy &x) { return as_impl<T>::from(x.get()); }
template <class T> T as(const std::string &x) { return T(x); }

// table() of a factor-coded integer vector: counts of the codes 1..max
inline IntegerVector table(const IntegerVector &codes) {
  int mx = 0;
  for (int v : codes) mx = std::max(mx, v);
  IntegerVector t(mx);
  for (int v : codes) if (v >= 1) t[v - 1] += 1;
  return t;
}

// rownames(x) / colnames(x): readable as a CharacterVector, assignable
struct DimNamesProxy {
  SEXP node; const char *which;
  operator CharacterVector() const {
    auto it = node->attrs.find(which);
    return it == node->attrs.end() ? CharacterVector() : CharacterVector(it->second);
  }
  DimNamesProxy &operator=(const CharacterVector &v) { node->attrs[which] = v.node; return *this; }
  DimNamesProxy &operator=(const DimNamesProxy &v) { node->attrs[which] = CharacterVector(v).node; return *this; }
};
inline DimNamesProxy rownames(const RObject &x) { return DimNamesProxy{x.node, "rownames"}; }
inline DimNamesProxy colnames(const RObject &x) { return DimNamesProxy{x.node, "colnames"}; }

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
