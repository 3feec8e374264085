// pt_device.cuh -- device-side arithmetic vocabulary of the path tracer.
//
// Parity rule (SURVEY.md section 7.3 / DESIGN.md "numerics"): every + - * /
// sqrt of the reference is reproduced as ONE correctly-rounded IEEE-754
// binary32 operation in the same association order; nothing is contracted
// into FMA except where the reference itself calls sycl::fma (vec.hpp:12).
// The helpers below use the explicit round-to-nearest intrinsics, which nvcc
// never fuses, so the property does not depend on -fmad=false (the build sets
// it anyway).  Transcendentals (sin cos asin atan2 log pow) follow glibc's own
// float algorithms operation by operation (pt_glibc_math.cuh), so the GPU
// render is the reference's CPU render bit for bit.
#ifndef PT_DEVICE_CUH
#define PT_DEVICE_CUH

#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <math_constants.h>
#define PT_STAT(counter)
#else
#include "pt_hostshim.h"  // tests/host: a CPU build of these headers, for the scan-equivalence tests only
#endif
#include <stdint.h>

namespace ptb {

#define PT_DEV __device__ __forceinline__

PT_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
PT_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
PT_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
// IEEE division and square root, inlined: a few independent ones in a row (a normal's three components, both
// roots of a sphere) then overlap, which matters in the short rounds that bound the frame's critical path
// (measured: 1.12 -> 1.21 Gpaths/s on the default scene).  -DPT_OUTLINE_DIV keeps one out-of-line copy instead.
#ifndef PT_OUTLINE_DIV
PT_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }
PT_DEV float fsqrt(float a) { return __fsqrt_rn(a); }
#else
static __device__ __noinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
static __device__ __noinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
#endif

struct V3 {
  float x, y, z;
};
PT_DEV V3 v3(float x, float y, float z) { return V3 { x, y, z }; }
PT_DEV V3 vld(const float* p) { return V3 { p[0], p[1], p[2] }; }
PT_DEV V3 vadd(V3 a, V3 b) { return v3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
PT_DEV V3 vsub(V3 a, V3 b) { return v3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
PT_DEV V3 vmul(V3 a, V3 b) { return v3(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
PT_DEV V3 vscale(float s, V3 a) { return v3(fmul(s, a.x), fmul(s, a.y), fmul(s, a.z)); }
PT_DEV V3 vdivs(V3 a, float s) { return v3(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }
// dot = (x*x' + y*y') + z*z'  (the oracle shim's definition of sycl::dot)
PT_DEV float vdot(V3 a, V3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
PT_DEV V3 vcross(V3 a, V3 b) {
  return v3(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
            fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
PT_DEV float vlength(V3 a) { return fsqrt(vdot(a, a)); }
PT_DEV V3 unit_vector(V3 v) { return vdivs(v, vlength(v)); }  // vec.hpp:23
// vec.hpp:11-13 -- the one place the reference asks for a fused multiply-add
PT_DEV float length_squared(V3 v) { return __fmaf_rn(v.x, v.x, __fmaf_rn(v.y, v.y, fmul(v.z, v.z))); }
// vec.hpp:26: v - 2*dot(v,n)*n
PT_DEV V3 reflect(V3 v, V3 n) { return vsub(v, vscale(fmul(2.0f, vdot(v, n)), n)); }
// vec.hpp:29-35
PT_DEV V3 refract(V3 uv, V3 n, float etai_over_etat) {
  const float cos_theta = fminf(-vdot(uv, n), 1.0f);
  const V3 r_out_perp = vscale(etai_over_etat, vadd(uv, vscale(cos_theta, n)));
  const V3 r_out_parallel = vscale(-fsqrt(fabsf(fsub(1.0f, length_squared(r_out_perp)))), n);
  return vadd(r_out_perp, r_out_parallel);
}

}  // namespace ptb
#include "pt_glibc_math.cuh"
namespace ptb {

// ---- transcendentals: sin, cos, log, pow(x, 5), asin and atan2 are glibc's own algorithms, bit for bit
// (pt_glibc_math.cuh): the reference's host calls sinf / cosf / logf / powf / asinf / atan2f, which are not correctly
// rounded, so "binary64 and round once" differs from them in the last bit now and then -- enough to flip a
// constant_medium's hit / pass decision or a dielectric's reflect / refract decision a few times per million.
// Out-of-line (one copy in the kernel image; the scan loop must own the instruction cache) and by value (nothing
// forced into local memory).
#ifndef PT_MATH_FN_SINCOS
#define PT_MATH_FN_SINCOS __noinline__
#endif
#ifndef PT_MATH_FN_POW
#define PT_MATH_FN_POW __noinline__
#endif
#ifndef PT_MATH_FN_REST
#define PT_MATH_FN_REST __noinline__
#endif
#ifndef PT_MATH_BINARY64  // (experiments: the round-1 evaluation of all of them)
static __device__ PT_MATH_FN_SINCOS float t_sin(float x) { return g_sinf(x); }
PT_DEV float t_cos(float x) { return g_cosf(x); }
static __device__ PT_MATH_FN_SINCOS float2 t_sincos2(float x) { return make_float2(g_sinf(x), g_cosf(x)); }
#else
static __device__ __noinline__ float t_sin(float x) { return __double2float_rn(sin((double)x)); }
PT_DEV float t_cos(float x) { return __double2float_rn(cos((double)x)); }
static __device__ __noinline__ float2 t_sincos2(float x) {
  double ds, dc;
  sincos((double)x, &ds, &dc);
  return make_float2(__double2float_rn(ds), __double2float_rn(dc));
}
#endif
PT_DEV void t_sincos(float x, float& s, float& c) {
  const float2 r = t_sincos2(x);
  s = r.x, c = r.y;
}
#ifndef PT_MATH_BINARY64
static __device__ PT_MATH_FN_REST float t_asin(float x) { return g_asinf(x); }
static __device__ PT_MATH_FN_REST float t_atan2(float y, float x) { return g_atan2f(y, x); }
#else
PT_DEV float t_asin(float x) { return __double2float_rn(asin((double)x)); }
PT_DEV float t_atan2(float y, float x) { return __double2float_rn(atan2((double)y, (double)x)); }
#endif
#ifndef PT_MATH_BINARY64
static __device__ PT_MATH_FN_REST float t_log(float x) { return g_logf(x); }
#else
static __device__ __noinline__ float t_log(float x) { return __double2float_rn(log((double)x)); }
#endif
// pow(x, 5.0f) (material.hpp:65)
#ifndef PT_MATH_BINARY64
static __device__ PT_MATH_FN_POW float t_pow5(float x) { return g_pow5(x); }
#else
PT_DEV float t_pow5(float x) {  // x^5 by binary64 products (4 roundings at 2^-53)
  const double d = (double)x;
  const double d2 = d * d;
  return __double2float_rn(d2 * d2 * d);
}
#endif
// fmod(x, 1.0f) (texture.hpp:140,143): exact, like every fmod
PT_DEV float t_fmod1(float x) { return fsub(x, truncf(x)); }

constexpr float kPi = 3.1415926535897932385f;  // rtweekend.hpp:22
#define PT_INF_F (__int_as_float(0x7f800000))
#define kInf PT_INF_F

// ---- RNG: xorshift.hpp:64-93 (<32>: 7,1,9), rtweekend.hpp:33-92 -----------
struct Rng {
  uint32_t s;
};
PT_DEV uint32_t rng_next(Rng& g) {
  uint32_t s = g.s;
  s ^= s >> 7;
  s ^= s << 1;
  s ^= s >> 9;
  g.s = s;
  return s;
}
// rtweekend.hpp:39-42: generator() * 2^-32; the u32 -> float conversion rounds to nearest
PT_DEV float rng_float(Rng& g) { return fmul(__uint2float_rn(rng_next(g)), 2.3283064365386963e-10f); }
// rtweekend.hpp:45-48
PT_DEV float rng_range(Rng& g, float lo, float hi) { return fadd(lo, fmul(fsub(hi, lo), rng_float(g))); }
// rtweekend.hpp:60-67
PT_DEV V3 rng_unit_vec(Rng& g) {
  const float x = rng_range(g, -1.f, 1.f);
  const float maxy = fsqrt(fsub(1.f, fmul(x, x)));
  const float y = rng_range(g, -maxy, maxy);
  const float absz = fsqrt(fsub(fmul(maxy, maxy), fmul(y, y)));
  const float z = (rng_float(g) > 0.5f) ? absz : -absz;
  return v3(x, y, z);
}
// rtweekend.hpp:70-80
PT_DEV V3 rng_in_unit_ball(Rng& g) {
  const float r = rng_float(g);
  const float theta = rng_range(g, 0.f, fmul(2.f, kPi));
  const float phi = rng_range(g, 0.f, kPi);
  float sp, cp, st, ct;
  t_sincos(phi, sp, cp);
  t_sincos(theta, st, ct);
  const float plan_seed = fmul(r, sp);
  const float z = fmul(r, cp);
  return v3(fmul(plan_seed, ct), fmul(plan_seed, st), z);
}
// rtweekend.hpp:83-88
PT_DEV void rng_in_unit_disk(Rng& g, float& x, float& y) {
  x = rng_range(g, -1.f, 1.f);
  const float maxy = fsqrt(fsub(1.f, fmul(x, x)));
  y = rng_range(g, -maxy, maxy);
}

struct Ray {
  V3 o, d;
  float tm;
};
PT_DEV V3 ray_at(const Ray& r, float t) { return vadd(r.o, vscale(t, r.d)); }  // ray.hpp:21

}  // namespace ptb
#endif
