// TEST INFRASTRUCTURE ONLY -- never linked into the product (libptb200.so).
//
// Minimal host stand-in for the SYCL subset that triSYCL/path_tracer touches,
// so that the reference headers under /root/reference/include can be compiled
// UNMODIFIED, in place, into oracle/_ref/libptref.so (see oracle/Makefile).
//
// triSYCL itself is an unpinned third-party dependency that is not present in
// the reference tree (README.md:60-64 of the reference), so the arithmetic at
// this boundary is *defined here* and parity is "unpinned" at exactly these
// points (DESIGN.md, section "Oracle"):
//   dot(a,b)   = (a.x*b.x + a.y*b.y) + a.z*b.z      (left-to-right inner product)
//   length(a)  = sqrtf(dot(a,a))
//   cross(a,b) = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
//   vec op vec / vec op scalar / scalar op vec = element-wise IEEE-754 binary32
//   sin cos tan asin atan2 log pow fmod fmin fabs fma = the <cmath> float overloads
// The library built from this must be compiled with -ffp-contract=off.
//
// Execution model: handler::parallel_for runs the kernel functor on the host,
// rows (outer range dimension) distributed over OpenMP threads with
// schedule(runtime) -- triSYCL's host device uses a plain `#pragma omp for`
// over the outer dimension.
#ifndef PT_ORACLE_SYCL_SHIM_HPP
#define PT_ORACLE_SYCL_SHIM_HPP

#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <tuple>
#include <type_traits>
#include <typeinfo>
#include <utility>
#include <vector>

namespace ptref_shim {
// Observation hooks used by oracle/capture_main.cpp (scene capture from the
// unmodified reference main.cpp).  All null by default.
struct hooks_t {
  // a buffer was constructed over caller memory
  void (*on_host_buffer)(const std::type_info& elem, const void* ptr, std::size_t count) = nullptr;
  // a kernel is about to be launched; return true to SKIP executing it
  bool (*on_parallel_for)(const void* closure, std::size_t closure_bytes, std::size_t rows,
                          std::size_t cols) = nullptr;
};
inline hooks_t& hooks() {
  static hooks_t h;
  return h;
}
}  // namespace ptref_shim

namespace sycl {

// ------------------------------------------------------------------ float3
// 12 bytes, like triSYCL's array-backed vec<float,3> (the reference's camera
// is 7 float3 + 3 float = 96 B and its framebuffer 12 B per pixel).
class float3 {
  float e_[3];

 public:
  constexpr float3() : e_ { 0.f, 0.f, 0.f } {}
  template <typename A, typename = std::enable_if_t<std::is_arithmetic_v<A>>>
  constexpr float3(A a)
      : e_ { static_cast<float>(a), static_cast<float>(a), static_cast<float>(a) } {}
  template <typename A, typename B, typename C,
            typename = std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> &&
                                        std::is_arithmetic_v<C>>>
  constexpr float3(A a, B b, C c)
      : e_ { static_cast<float>(a), static_cast<float>(b), static_cast<float>(c) } {}

  constexpr float x() const { return e_[0]; }
  constexpr float y() const { return e_[1]; }
  constexpr float z() const { return e_[2]; }
  float& x() { return e_[0]; }
  float& y() { return e_[1]; }
  float& z() { return e_[2]; }

#define PTREF_F3_COMPOUND(OP)                                           \
  float3& operator OP##=(const float3& o) {                             \
    e_[0] OP## = o.e_[0];                                               \
    e_[1] OP## = o.e_[1];                                               \
    e_[2] OP## = o.e_[2];                                               \
    return *this;                                                       \
  }                                                                     \
  template <typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>> \
  float3& operator OP##=(S s) {                                         \
    const float f = static_cast<float>(s);                              \
    e_[0] OP## = f;                                                     \
    e_[1] OP## = f;                                                     \
    e_[2] OP## = f;                                                     \
    return *this;                                                       \
  }
  PTREF_F3_COMPOUND(+)
  PTREF_F3_COMPOUND(-)
  PTREF_F3_COMPOUND(*)
  PTREF_F3_COMPOUND(/)
#undef PTREF_F3_COMPOUND
};

#define PTREF_F3_BINARY(OP)                                                               \
  inline float3 operator OP(const float3& a, const float3& b) {                           \
    return float3(a.x() OP b.x(), a.y() OP b.y(), a.z() OP b.z());                        \
  }                                                                                       \
  template <typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>             \
  inline float3 operator OP(const float3& a, S s) {                                       \
    const float f = static_cast<float>(s);                                                \
    return float3(a.x() OP f, a.y() OP f, a.z() OP f);                                    \
  }                                                                                       \
  template <typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>             \
  inline float3 operator OP(S s, const float3& b) {                                       \
    const float f = static_cast<float>(s);                                                \
    return float3(f OP b.x(), f OP b.y(), f OP b.z());                                    \
  }
PTREF_F3_BINARY(+)
PTREF_F3_BINARY(-)
PTREF_F3_BINARY(*)
PTREF_F3_BINARY(/)
#undef PTREF_F3_BINARY
// NB: no unary minus here -- the reference defines it itself (vec.hpp:20).

// ------------------------------------------------------------------ math
inline float dot(const float3& a, const float3& b) {
  return (a.x() * b.x() + a.y() * b.y()) + a.z() * b.z();
}
inline float3 cross(const float3& a, const float3& b) {
  return float3(a.y() * b.z() - a.z() * b.y(), a.z() * b.x() - a.x() * b.z(),
                a.x() * b.y() - a.y() * b.x());
}
inline float sqrt(float v) { return std::sqrt(v); }
inline float length(const float3& a) { return std::sqrt(dot(a, a)); }
inline float sin(float v) { return std::sin(v); }
inline float cos(float v) { return std::cos(v); }
inline float tan(float v) { return std::tan(v); }
inline float asin(float v) { return std::asin(v); }
inline float atan2(float a, float b) { return std::atan2(a, b); }
inline float log(float v) { return std::log(v); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float fmod(float a, float b) { return std::fmod(a, b); }
inline float fmin(float a, float b) { return std::fmin(a, b); }
inline float fabs(float v) { return std::fabs(v); }
inline float fma(float a, float b, float c) { return std::fma(a, b, c); }

// ------------------------------------------------------------------ ranges
template <int N> class range {
  std::size_t d_[N];

 public:
  range() : d_ {} {}
  template <int M = N, typename = std::enable_if_t<M == 1>>
  range(std::size_t a) : d_ { a } {}
  template <int M = N, typename = std::enable_if_t<M == 2>>
  range(std::size_t a, std::size_t b) : d_ { a, b } {}
  std::size_t operator[](int i) const { return d_[i]; }
  std::size_t size() const {
    std::size_t s = 1;
    for (int i = 0; i < N; ++i) s *= d_[i];
    return s;
  }
};

template <int N> class id {
  std::size_t d_[N];

 public:
  id() : d_ {} {}
  template <int M = N, typename = std::enable_if_t<M == 2>>
  id(std::size_t a, std::size_t b) : d_ { a, b } {}
  std::size_t operator[](int i) const { return d_[i]; }
};

template <int N> class item {
  id<N> id_;
  range<N> range_;

 public:
  item(const id<N>& i, const range<N>& r) : id_ { i }, range_ { r } {}
  id<N> get_id() const { return id_; }
  range<N> get_range() const { return range_; }
  std::size_t get_linear_id() const {
    static_assert(N == 2);
    return id_[0] * range_[1] + id_[1];
  }
};

namespace access {
enum class mode { read, write, read_write, discard_write, discard_read_write };
}

template <typename T> class global_ptr {
  T* p_ = nullptr;

 public:
  global_ptr() = default;
  global_ptr(T* p) : p_ { p } {}
  T& operator[](std::size_t i) const { return p_[i]; }
  T* get() const { return p_; }
};

class handler;

// ------------------------------------------------------------------ buffer / accessor
template <typename T, int N> class accessor {
  T* data_;
  range<N> range_;

  struct row_proxy {
    T* row;
    T& operator[](std::size_t c) const { return row[c]; }
  };

 public:
  accessor(T* d, const range<N>& r) : data_ { d }, range_ { r } {}
  std::size_t get_count() const { return range_.size(); }
  range<N> get_range() const { return range_; }
  global_ptr<T> get_pointer() const { return global_ptr<T> { data_ }; }
  auto operator[](std::size_t i) const -> std::conditional_t<N == 1, T&, row_proxy> {
    if constexpr (N == 1)
      return data_[i];
    else
      return row_proxy { data_ + i * range_[1] };
  }
};

template <typename T, int N> class buffer {
  std::shared_ptr<std::vector<T>> owned_;
  T* data_ = nullptr;
  range<N> range_;

 public:
  buffer(const range<N>& r) : owned_ { std::make_shared<std::vector<T>>(r.size()) }, range_ { r } {
    data_ = owned_->data();
  }
  buffer(T* host, const range<N>& r) : data_ { host }, range_ { r } {
    if (auto h = ptref_shim::hooks().on_host_buffer) h(typeid(T), host, r.size());
  }
  range<N> get_range() const { return range_; }
  T* host_data() const { return data_; }
  template <access::mode M> accessor<T, N> get_access(handler&) { return { data_, range_ }; }
  template <access::mode M> accessor<T, N> get_access() { return { data_, range_ }; }
};

// ------------------------------------------------------------------ handler / queue
class handler {
 public:
  template <typename Name = void, typename K> void single_task(K&& k) { k(); }

  template <typename Name = void, typename K> void parallel_for(range<2> r, K&& k) {
    if (auto h = ptref_shim::hooks().on_parallel_for)
      if (h(&k, sizeof(k), r[0], r[1])) return;
    const long rows = static_cast<long>(r[0]);
    const long cols = static_cast<long>(r[1]);
#pragma omp parallel for schedule(runtime)
    for (long y = 0; y < rows; ++y)
      for (long x = 0; x < cols; ++x)
        k(item<2> { id<2> { static_cast<std::size_t>(y), static_cast<std::size_t>(x) }, r });
  }
};

class queue {
 public:
  template <typename F> void submit(F&& f) {
    handler h;
    f(h);
  }
  void wait() {}
};

}  // namespace sycl

#endif
