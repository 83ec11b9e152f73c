// Stand-in for <boost/optional.hpp>: the subset used by the reference (has_value/value/none).
#pragma once
#include <stdexcept>
namespace boost {
struct none_t {};
static const none_t none = none_t();
template <class T> class optional {
  bool has_ = false;
  alignas(T) unsigned char buf_[sizeof(T)];
public:
  optional() {}
  optional(none_t) {}
  optional(const T &v) : has_(true) { new (buf_) T(v); }
  optional(const optional &o) : has_(o.has_) { if (has_) new (buf_) T(o.value()); }
  optional &operator=(const optional &o) {
    if (this != &o) { reset(); has_ = o.has_; if (has_) new (buf_) T(o.value()); }
    return *this;
  }
  optional &operator=(const T &v) { reset(); has_ = true; new (buf_) T(v); return *this; }
  ~optional() { reset(); }
  void reset() { if (has_) { reinterpret_cast<T *>(buf_)->~T(); has_ = false; } }
  bool has_value() const { return has_; }
  const T &value() const {
    if (!has_) throw std::logic_error("optional: no value");
    return *reinterpret_cast<const T *>(buf_);
  }
};
} // namespace boost
