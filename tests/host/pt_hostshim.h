// TEST INFRASTRUCTURE ONLY -- lets the parity-critical device headers (pt_device.cuh, pt_prims.cuh) be
// compiled by g++ so that the closest-hit scan with and without culling can be compared on the CPU, ray by
// ray, without a GPU (tests/host/scan_check.cpp, tests/test_flat_culling.py).  Every CUDA intrinsic the
// headers use is restated with the same IEEE-754 semantics; compile with -ffp-contract=off.  The product
// never includes this file: it is selected by the absence of __CUDACC__.
#ifndef PT_HOSTSHIM_H
#define PT_HOSTSHIM_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __global__
#ifndef __restrict__
#define __restrict__
#endif

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2 { x, y }; }
inline float4 make_float4(float x, float y, float z, float w) { return float4 { x, y, z, w }; }

struct PtHostDim { unsigned x, y, z; };
static PtHostDim threadIdx = { 0, 0, 0 };

using std::max;
using std::min;

inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __double2float_rn(double a) { return (float)a; }
inline float __uint2float_rn(unsigned a) { return (float)a; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
  const unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (unsigned)((v << (shift & 31)) >> 32);
}
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }  // a team of one
inline unsigned __ballot_sync(unsigned, bool pred) { return pred ? 1u : 0u; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline void sincos(double x, double* s, double* c) { *s = sin(x), *c = cos(x); }

// work counters of the host build (PT_STAT is empty in the device build)
struct PtHostStats { unsigned long long flat_nodes, graze_nodes, triangle_tests, graze_tests; };
static thread_local PtHostStats pt_host_stats = { 0, 0, 0, 0 };
#define PT_STAT(counter) (++pt_host_stats.counter)
#endif
