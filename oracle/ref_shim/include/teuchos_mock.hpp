// Minimal stand-ins for the few Teuchos facilities that the hot-path routines of FEDDLib's FE_def.hpp touch.
// Test infrastructure only (oracle/_ref): lets the reference's own assembly loops compile unmodified
// without Trilinos.  Nothing here is used by the product.
#pragma once
#include <cstddef>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <sstream>
#include <string>
#include <vector>

#define TEUCHOS_TEST_FOR_EXCEPTION(cond, exc, msg)                 \
    do {                                                            \
        if (cond) {                                                 \
            std::ostringstream teuchos_os__;                        \
            teuchos_os__ << msg;                                    \
            throw exc(teuchos_os__.str());                          \
        }                                                           \
    } while (0)

namespace Teuchos {

struct ENull { };
static const ENull null = ENull();

template <class T>
class RCP {
  public:
    RCP() {}
    RCP(ENull) {}
    explicit RCP(T *p) : p_(p) {}
    RCP(const std::shared_ptr<T> &p) : p_(p) {}
    template <class U> RCP(const RCP<U> &o) : p_(o.shared()) {}
    T *operator->() const { return p_.get(); }
    T &operator*() const { return *p_; }
    T *get() const { return p_.get(); }
    bool is_null() const { return !p_; }
    void reset() { p_.reset(); }
    void reset(T *p) { p_.reset(p); }
    const std::shared_ptr<T> &shared() const { return p_; }
  private:
    std::shared_ptr<T> p_;
};

template <class T> RCP<T> rcp(T *p) { return RCP<T>(p); }
template <class T, class U> RCP<T> rcp_const_cast(const RCP<U> &p) { return RCP<T>(std::const_pointer_cast<T>(p.shared())); }
template <class T, class U> RCP<T> rcp_dynamic_cast(const RCP<U> &p) { return RCP<T>(std::dynamic_pointer_cast<T>(p.shared())); }

template <class T>
class ArrayView {
  public:
    ArrayView() : p_(nullptr), n_(0) {}
    ArrayView(const T *p, std::size_t n) : p_(p), n_(n) {}
    const ArrayView &operator()() const { return *this; }
    std::size_t size() const { return n_; }
    const T &operator[](std::size_t i) const { return p_[i]; }
    const T *getRawPtr() const { return p_; }
  private:
    const T *p_;
    std::size_t n_;
};

template <class T>
class Array {
  public:
    Array() {}
    explicit Array(std::size_t n, const T &v = T()) : v_(n, v) {}
    std::size_t size() const { return v_.size(); }
    T &operator[](std::size_t i) { return v_[i]; }
    const T &operator[](std::size_t i) const { return v_[i]; }
    ArrayView<T> operator()() const { return ArrayView<T>(v_.data(), v_.size()); }
    void push_back(const T &x) { v_.push_back(x); }
    void resize(std::size_t n) { v_.resize(n); }
  private:
    std::vector<T> v_;
};

template <class T>
class ArrayRCP {
  public:
    ArrayRCP() : p_(nullptr), n_(0) {}
    ArrayRCP(ENull) : p_(nullptr), n_(0) {}
    ArrayRCP(T *p, std::size_t n) : p_(p), n_(n) {}
    T &operator[](std::size_t i) const { return p_[i]; }
    std::size_t size() const { return n_; }
  private:
    T *p_;
    std::size_t n_;
};

template <class T> struct ScalarTraits { static T eps() { return std::numeric_limits<T>::epsilon(); } static T zero() { return T(0); } static T one() { return T(1); } };
template <class T> struct OrdinalTraits { static T invalid() { return T(-1); } };

class ParameterList { };
class Time { };
class TimeMonitor {
  public:
    TimeMonitor() {}
    explicit TimeMonitor(Time &) {}
    static RCP<Time> getNewCounter(const std::string &) { return RCP<Time>(new Time()); }
};
class CommandLineProcessor { };
template <class O, class S> class BLAS { };

} // namespace Teuchos
