// TEST INFRASTRUCTURE ONLY -- the product's restatement of glibc's sinf / cosf / logf (pt_glibc_math.cuh, compiled for
// the host through pt_hostshim.h) against the running libm, on EVERY binary32 argument.  tests/test_glibc_math.py.
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>

#include "pt_device.cuh"

using namespace ptb;

extern "C" {
// kind: 0 sinf, 1 cosf, 2 logf, 3 powf(x, 5), 4 asinf, 5 atanf.  Checks the arguments with bit patterns [first, last] step `stride`; returns the number
// of mismatches (NaN == NaN) and the first mismatching argument's bits in *bad.
uint64_t math_check(int kind, uint64_t first, uint64_t last, uint64_t stride, uint32_t* bad) {
  uint64_t mismatches = 0;
  uint32_t first_bad = 0;
  bool have = false;
#pragma omp parallel for schedule(static, 1 << 16) reduction(+ : mismatches)
  for (long long b = (long long)first; b <= (long long)last; b += (long long)stride) {
    const uint32_t bits = (uint32_t)b;
    float x;
    std::memcpy(&x, &bits, 4);
    const float want = kind == 0 ? sinf(x) : kind == 1 ? cosf(x) : kind == 2 ? logf(x) : kind == 3 ? powf(x, 5.0f) : kind == 4 ? asinf(x) : atanf(x);
    const float got = kind == 0 ? g_sinf(x) : kind == 1 ? g_cosf(x) : kind == 2 ? g_logf(x) : kind == 3 ? g_pow5(x) : kind == 4 ? g_asinf(x) : g_atanf(x);
    uint32_t wb, gb;
    std::memcpy(&wb, &want, 4), std::memcpy(&gb, &got, 4);
    if (wb != gb && !(want != want && got != got)) {
      ++mismatches;
#pragma omp critical
      if (!have) have = true, first_bad = bits;
    }
  }
  *bad = first_bad;
  return mismatches;
}
// The values themselves (for the GPU comparison): out[i] = f(in[i]) by the running libm.
// atan2f on pairs: n_pairs pseudo-random (y, x) bit patterns from `seed`, a quarter of them of comparable magnitude, a
// quarter on the unit circle (what the renderer passes), plus every pair of a list of special values.
uint64_t math_check_atan2(uint64_t seed, uint64_t n_pairs, uint32_t* bad_y, uint32_t* bad_x) {
  uint64_t mismatches = 0;
  bool have = false;
  const float special[] = { 0.f, -0.f, 1.f, -1.f, INFINITY, -INFINITY, NAN, 1e-45f, -1e-45f, 1e-38f, 3e38f, -3e38f, 0.5f, 2.f, 1e-20f, 1e20f };
  const long long n_special = 16 * 16;
#pragma omp parallel for schedule(static, 1 << 14) reduction(+ : mismatches)
  for (long long i = 0; i < (long long)n_pairs + n_special; ++i) {
    float y, x;
    if (i < n_special) {
      y = special[i / 16], x = special[i % 16];
    } else {
      uint64_t s = seed * 0x9e3779b97f4a7c15ull + (uint64_t)i * 0xbf58476d1ce4e5b9ull + 1ull;
      s ^= s >> 12, s ^= s << 25, s ^= s >> 27, s *= 2685821657736338717ull;
      uint32_t by = (uint32_t)s, bx = (uint32_t)(s >> 32);
      std::memcpy(&y, &by, 4), std::memcpy(&x, &bx, 4);
      if ((i & 3) == 1) {  // comparable magnitudes
        bx = (bx & 0x807fffffu) | (by & 0x7f800000u);
        std::memcpy(&x, &bx, 4);
      } else if ((i & 3) == 2) {  // a point of the unit circle
        const double a = (double)(s >> 11) * (6.283185307179586 / 9007199254740992.0);
        y = (float)sin(a), x = (float)cos(a);
      }
    }
    const float want = atan2f(y, x), got = g_atan2f(y, x);
    uint32_t wb, gb;
    std::memcpy(&wb, &want, 4), std::memcpy(&gb, &got, 4);
    if (wb != gb && !(want != want && got != got)) {
      ++mismatches;
#pragma omp critical
      if (!have) have = true, std::memcpy(bad_y, &y, 4), std::memcpy(bad_x, &x, 4);
    }
  }
  return mismatches;
}
void math_libm(int kind, int n, const float* in, float* out) {
  if (kind == 6) {
    for (int i = 0; i < n; ++i) out[i] = atan2f(in[2 * i], in[2 * i + 1]);
    return;
  }
  if (kind >= 4) {
    for (int i = 0; i < n; ++i) out[i] = kind == 4 ? asinf(in[i]) : atanf(in[i]);
    return;
  }
  for (int i = 0; i < n; ++i) out[i] = kind == 0 ? sinf(in[i]) : kind == 1 ? cosf(in[i]) : kind == 2 ? logf(in[i]) : powf(in[i], 5.0f);
}
}
