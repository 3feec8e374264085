// pt/sycl_facade.hpp -- the sliver of the SYCL host API that scene scripts written for
// triSYCL/path_tracer (its src/main.cpp) touch, so they compile unchanged against the B200
// library.  Nothing here executes kernels: rendering goes through the C-ABI (pt_abi.h).
//
//   sycl::float3            12-byte value type, element-wise IEEE binary32 arithmetic
//   sycl::dot/cross/length/sqrt/tan ...   host math used while building scenes and cameras
//   sycl::range / buffer / accessor / queue / access::mode   storage for the framebuffer
//
// Arithmetic contract (must equal what the reference's host code computes, because the scene a
// script builds is part of the parity input): dot = (x*x' + y*y') + z*z', length = sqrtf(dot),
// textbook cross product, all other operators element-wise, no contraction (build the host code
// with -ffp-contract=off).
#ifndef PT_SYCL_FACADE_HPP
#define PT_SYCL_FACADE_HPP

#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <tuple>
#include <type_traits>
#include <vector>

namespace sycl {

template <typename T> inline constexpr bool pt_is_number = std::is_arithmetic_v<std::remove_cv_t<std::remove_reference_t<T>>>;

class float3 {
 public:
  constexpr float3() = default;
  template <typename N, std::enable_if_t<pt_is_number<N>, int> = 0>
  constexpr float3(N all) : v_ { float(all), float(all), float(all) } {}
  template <typename A, typename B, typename C,
            std::enable_if_t<pt_is_number<A> && pt_is_number<B> && pt_is_number<C>, int> = 0>
  constexpr float3(A a, B b, C c) : v_ { float(a), float(b), float(c) } {}

  constexpr float x() const { return v_[0]; }
  constexpr float y() const { return v_[1]; }
  constexpr float z() const { return v_[2]; }
  constexpr float& x() { return v_[0]; }
  constexpr float& y() { return v_[1]; }
  constexpr float& z() { return v_[2]; }
  constexpr float operator[](int i) const { return v_[i]; }

  template <typename F> constexpr float3& apply(const float3& o, F f) {
    for (int i = 0; i < 3; ++i) v_[i] = f(v_[i], o.v_[i]);
    return *this;
  }
  constexpr float3& operator+=(const float3& o) { return apply(o, [](float a, float b) { return a + b; }); }
  constexpr float3& operator-=(const float3& o) { return apply(o, [](float a, float b) { return a - b; }); }
  constexpr float3& operator*=(const float3& o) { return apply(o, [](float a, float b) { return a * b; }); }
  constexpr float3& operator/=(const float3& o) { return apply(o, [](float a, float b) { return a / b; }); }

 private:
  float v_[3] = { 0.f, 0.f, 0.f };
};
static_assert(sizeof(float3) == 12);

// vector (op) vector, vector (op) scalar, scalar (op) vector: element-wise, scalars converted to float first
#define PT_FLOAT3_OP(OP)                                                                         \
  constexpr float3 operator OP(const float3& a, const float3& b) {                               \
    return { a.x() OP b.x(), a.y() OP b.y(), a.z() OP b.z() };                                   \
  }                                                                                              \
  template <typename N, std::enable_if_t<pt_is_number<N>, int> = 0>                              \
  constexpr float3 operator OP(const float3& a, N s) {                                           \
    return { a.x() OP float(s), a.y() OP float(s), a.z() OP float(s) };                          \
  }                                                                                              \
  template <typename N, std::enable_if_t<pt_is_number<N>, int> = 0>                              \
  constexpr float3 operator OP(N s, const float3& b) {                                           \
    return { float(s) OP b.x(), float(s) OP b.y(), float(s) OP b.z() };                          \
  }
PT_FLOAT3_OP(+)
PT_FLOAT3_OP(-)
PT_FLOAT3_OP(*)
PT_FLOAT3_OP(/)
#undef PT_FLOAT3_OP
// (unary minus is deliberately absent, as in SYCL 1.2.1; scene scripts bring their own)

inline float dot(const float3& a, const float3& b) { return (a.x() * b.x() + a.y() * b.y()) + a.z() * b.z(); }
inline float3 cross(const float3& a, const float3& b) {
  return { a.y() * b.z() - a.z() * b.y(), a.z() * b.x() - a.x() * b.z(), a.x() * b.y() - a.y() * b.x() };
}
inline float length(const float3& a) { return std::sqrt(dot(a, a)); }
inline float sqrt(float v) { return std::sqrt(v); }
inline float tan(float v) { return std::tan(v); }
inline float sin(float v) { return std::sin(v); }
inline float cos(float v) { return std::cos(v); }
inline float fabs(float v) { return std::fabs(v); }
inline float fmin(float a, float b) { return std::fmin(a, b); }
inline float pow(float a, float b) { return std::pow(a, b); }

template <int N> class range {
 public:
  range() = default;
  template <typename... S, std::enable_if_t<sizeof...(S) == N, int> = 0>
  range(S... s) : n_ { std::size_t(s)... } {}
  std::size_t operator[](int i) const { return n_[i]; }
  std::size_t size() const {
    std::size_t t = 1;
    for (auto d : n_) t *= d;
    return t;
  }

 private:
  std::size_t n_[N] = {};
};

namespace access {
enum class mode { read, write, read_write, discard_write, discard_read_write };
}

// Host-side view of a buffer: a[i] for one dimension, a[row][col] for two.
template <typename T, int N> class accessor {
 public:
  accessor(T* base, range<N> r) : base_ { base }, r_ { r } {}
  std::size_t get_count() const { return r_.size(); }
  decltype(auto) operator[](std::size_t i) const {
    if constexpr (N == 1) {
      return (base_[i]);
    } else {
      struct row_view {
        T* p;
        T& operator[](std::size_t c) const { return p[c]; }
      };
      return row_view { base_ + i * r_[1] };
    }
  }
  T* get_pointer() const { return base_; }

 private:
  T* base_;
  range<N> r_;
};

template <typename T, int N> class buffer {
 public:
  explicit buffer(range<N> r) : own_ { std::make_shared<std::vector<T>>(r.size()) }, data_ { own_->data() }, r_ { r } {}
  buffer(T* host, range<N> r) : data_ { host }, r_ { r } {}
  range<N> get_range() const { return r_; }
  T* data() const { return data_; }
  template <access::mode> accessor<T, N> get_access() { return { data_, r_ }; }
  template <access::mode, typename H> accessor<T, N> get_access(H&) { return { data_, r_ }; }

 private:
  std::shared_ptr<std::vector<T>> own_;
  T* data_;
  range<N> r_;
};

// The B200 library is driven through a blocking C call; the queue only keeps scripts compiling.
class queue {
 public:
  void wait() {}
};

}  // namespace sycl
#endif
