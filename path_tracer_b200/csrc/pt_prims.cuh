// pt_prims.cuh -- the closest-hit scan of render.hpp:30-51: primitives (sphere / rect / triangle / box /
// constant_medium, each in the reference's operation order), the conservative sphere miss filter, chunk
// culling, and the per-ray scan drivers shared by both kernels.  Parity-critical: everything here is
// exercised bit for bit against the oracle; the schedulers (pt_wave.cu, pt_lane.cu) only decide which
// lane runs which piece.
#ifndef PT_PRIMS_CUH
#define PT_PRIMS_CUH
#include <stdint.h>

#include "pt_abi.h"
#include "pt_device.cuh"
#include "pt_packed.h"
#include "pt_stage.cuh"

namespace ptb {
namespace {

constexpr float kTMin = 0.001f;  // render.hpp:40
#ifndef PT_SCAN_UNROLL
#define PT_SCAN_UNROLL 8
#endif
constexpr int kScanUnroll = PT_SCAN_UNROLL;  // spheres per hot-loop trip
#ifndef PT_DEEP_RATE
#define PT_DEEP_RATE 20
#endif
constexpr int kDeepRate = PT_DEEP_RATE;  // lane kernel: a pixel is DEEP above kDeepBase + kDeepRate * samples scans so far
constexpr int kDeepBase = 64;
#ifndef PT_MIN_PIXELS_PER_TEAM
#define PT_MIN_PIXELS_PER_TEAM 3
#endif
constexpr int kMinPixelsPerTeam = PT_MIN_PIXELS_PER_TEAM;  // launch with larger teams below this many pixels per team

struct Best {
  float t;
  int id;
};

// Tie-break keys of every object (pt_packed.h); small enough to travel by value into out-of-line code.
struct KeyTable {
  const int32_t* keys;
  uint32_t base[6];
};
PT_DEV KeyTable key_table(const SceneDesc& sc) {
  KeyTable k;
  k.keys = sc.keys;
#pragma unroll
  for (int i = 0; i < 6; ++i) k.base[i] = sc.key_base[i];
  return k;
}
PT_DEV int key_of(const KeyTable& kt, int id) {
  const int type = id >> kIdShift;
  uint32_t b = kt.base[0];
#pragma unroll
  for (int i = 1; i < 6; ++i)
    if (type == i) b = kt.base[i];
  return kt.keys[b + (uint32_t)(id & (int)kIdMask)];
}
PT_DEV int key_of(const SceneDesc& sc, int id) {
  return sc.keys[sc.key_base[id >> kIdShift] + (uint32_t)(id & (int)kIdMask)];
}

// Winner rule: minimum t, then maximum key (pt_packed.h).  Called with a
// candidate that already satisfies its own primitive's range test.
template <typename Keys> PT_DEV void consider(const Keys& sc, Best& best, float t, int id) {
  if (t < best.t) {
    best.t = t, best.id = id;
  } else if (t == best.t) {
    if (best.id < 0 || key_of(sc, id) > key_of(sc, best.id)) best.t = t, best.id = id;
  }
}
// rect / triangle / box accept with `!(t > max)`, which lets NaN through
// (rectangle.hpp:36, triangle.hpp:91): mirror that.
template <typename Keys> PT_DEV void consider_le(const Keys& sc, Best& best, float t, int id) {
  if (t == best.t) {
    if (best.id < 0 || key_of(sc, id) > key_of(sc, best.id)) best.id = id;
  } else {
    best.t = t, best.id = id;
  }
}

// ---------------------------------------------------------------- primitives
// Exact roots of one sphere for the scan (sphere.hpp:74-105 with max = +inf;
// the running-closest filter is applied by consider()).  `a` = dot(d,d).
template <typename Keys>
PT_DEV void sphere_roots_scan(const Keys& sc, Best& best, const Ray& r, float a, float cx, float cy, float cz, float r2,
                              int id) {
  const float ocx = fsub(r.o.x, cx), ocy = fsub(r.o.y, cy), ocz = fsub(r.o.z, cz);
  const float b = fadd(fadd(fmul(ocx, r.d.x), fmul(ocy, r.d.y)), fmul(ocz, r.d.z));
  const float c = fsub(fadd(fadd(fmul(ocx, ocx), fmul(ocy, ocy)), fmul(ocz, ocz)), r2);
  const float disc = fsub(fmul(b, b), fmul(a, c));
  if (!(disc > 0.f)) return;
  // Both roots are <= 0 < t_min when the centre is behind an outside origin:
  // c > 0 gives sqrt(disc) <= b, so (-b + sqrt(disc))/a <= 0.
  if (b > 0.f && c > 0.f) return;
  const float sq = fsqrt(disc);
  const float t0 = fdiv(fsub(-b, sq), a);
  if (t0 < kInf && t0 > kTMin) {
    consider(sc, best, t0, id);
    return;
  }
  const float t1 = fdiv(fadd(-b, sq), a);
  if (t1 < kInf && t1 > kTMin) consider(sc, best, t1, id);
}

// sphere.hpp:59-106 in full, for constant_medium boundaries (arbitrary min/max).
PT_DEV bool sphere_hit_t(const Ray& r, V3 center, float r2, float tmin, float tmax, float& t_out) {
  const V3 oc = vsub(r.o, center);
  const float a = vdot(r.d, r.d);
  const float b = vdot(oc, r.d);
  const float c = fsub(vdot(oc, oc), r2);
  const float disc = fsub(fmul(b, b), fmul(a, c));
  if (disc > 0.f) {
    const float sq = fsqrt(disc);
    float temp = fdiv(fsub(-b, sq), a);
    if (temp < tmax && temp > tmin) {
      t_out = temp;
      return true;
    }
    temp = fdiv(fadd(-b, sq), a);
    if (temp < tmax && temp > tmin) {
      t_out = temp;
      return true;
    }
  }
  return false;
}

struct AxisSel {
  float ok, dk, oa, da, ob, db;
};
PT_DEV AxisSel axis_select(const Ray& r, int axis) {
  if (axis == PT_AXIS_XY) return AxisSel { r.o.z, r.d.z, r.o.x, r.d.x, r.o.y, r.d.y };
  if (axis == PT_AXIS_XZ) return AxisSel { r.o.y, r.d.y, r.o.x, r.d.x, r.o.z, r.d.z };
  return AxisSel { r.o.x, r.d.x, r.o.y, r.d.y, r.o.z, r.d.z };
}

// rectangle.hpp:31-49 / 69-87 / 107-125: returns hit and t (a, b = in-plane coordinates).
PT_DEV bool rect_hit_t(const Ray& r, int axis, float a0, float a1, float b0, float b1, float k, float tmin,
                       float tmax, float& t_out, float& a_out, float& b_out) {
  const AxisSel s = axis_select(r, axis);
  const float t = fdiv(fsub(k, s.ok), s.dk);
  if (t < tmin || t > tmax) return false;
  const float a = fadd(s.oa, fmul(t, s.da));
  const float b = fadd(s.ob, fmul(t, s.db));
  if (a < a0 || a > a1 || b < b0 || b > b1) return false;
  t_out = t, a_out = a, b_out = b;
  return true;
}

// box.hpp:29-50 over the six sides of box.hpp:20-25 (xy@p1.z, xy@p0.z, xz@p1.y, xz@p0.y, yz@p1.x,
// yz@p0.x; a later side takes a tie).  Returns the winning side or -1.  One loop body instead of six
// inlined rectangles keeps the code small.
PT_DEV bool box_side_hit_t(const Ray& r, V3 p0, V3 p1, int s, float tmin, float tmax, float& t, float& a, float& b) {
  const int axis = s >> 1;  // PT_AXIS_XY, PT_AXIS_XZ, PT_AXIS_YZ
  const bool hi = (s & 1) == 0;
  float a0, a1, b0, b1, k;
  if (axis == PT_AXIS_XY)
    a0 = p0.x, a1 = p1.x, b0 = p0.y, b1 = p1.y, k = hi ? p1.z : p0.z;
  else if (axis == PT_AXIS_XZ)
    a0 = p0.x, a1 = p1.x, b0 = p0.z, b1 = p1.z, k = hi ? p1.y : p0.y;
  else
    a0 = p0.y, a1 = p1.y, b0 = p0.z, b1 = p1.z, k = hi ? p1.x : p0.x;
  return rect_hit_t(r, axis, a0, a1, b0, b1, k, tmin, tmax, t, a, b);
}
PT_DEV int box_hit_t(const Ray& r, V3 p0, V3 p1, float tmin, float tmax, float& t_out, float& a_out,
                     float& b_out) {
  int side = -1;
  float closest = tmax;
#pragma unroll 1
  for (int s = 0; s < 6; ++s) {
    float t, a, b;
    if (box_side_hit_t(r, p0, p1, s, tmin, closest, t, a, b)) side = s, closest = t, t_out = t, a_out = a, b_out = b;
  }
  return side;
}

// triangle.hpp:58-100 (Moller-Trumbore) up to the range test; e1, e2 hoisted.
PT_DEV bool triangle_hit_t(const Ray& r, V3 v0, V3 e1, V3 e2, float tmin, float tmax, float& t_out) {
  const V3 h = vcross(r.d, e2);
  const float a = vdot(e1, h);
  const float a_abs = fabsf(a);
  if (a_abs < 0.0000001f) return false;
  const bool a_pos = a > 0.f;
  const V3 s = vsub(r.o, v0);
  const float u = vdot(s, h);
  const bool u_pos = u > 0.f;
  if ((u_pos != a_pos) || fabsf(u) > a_abs) return false;
  const V3 q = vcross(s, e1);
  const float v = vdot(r.d, q);
  const bool v_pos = v > 0.f;
  if ((v_pos != a_pos) || (fabsf(fadd(u, v)) > a_abs)) return false;
  const float length = fdiv(vdot(e2, q), a);
  if (length < tmin || length > tmax) return false;
  t_out = length;
  return true;
}

PT_DEV V3 moving_center(V3 c0, V3 dv, float f) { return vadd(c0, vscale(f, dv)); }  // sphere.hpp:55

// (ranges and helpers of the slab tests; "chunk culling" below explains them)
constexpr float kCullInvMax = 1.152921504606847e18f;  // 2^60
constexpr float kCullDirMin = 9.5367431640625e-7f;    // 2^-20
constexpr float kCullDirMax = 1048576.f;              // 2^20
PT_DEV float rcp_fast(float x) {
#ifdef __CUDACC__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
// Can NO side of the box p0..p1 be hit anywhere on the ray's line (box.hpp:29-50 with min = -inf, max = +inf)?  A side's
// hit passes the reference's own bounds check (rectangle.hpp:38-41), so its point lies on the box up to the rounding of
// o + t d: the line then crosses the box grown by any margin above that rounding -- here a generous 1e-3 of the distances
// involved.  true: provably no hit (the slab test of the grown box fails on the whole line); false: don't know.  Rays
// outside the range the clamped reciprocal is good for answer "don't know".
PT_DEV bool box_line_missed(const Ray& r, V3 p0, V3 p1) {
  const float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
  const float dmax = fmaxf(fmaxf(ax, ay), az), dmin = fminf(fminf(ax, ay), az);
  const float lox = fminf(p0.x, p1.x), hix = fmaxf(p0.x, p1.x), loy = fminf(p0.y, p1.y), hiy = fmaxf(p0.y, p1.y);
  const float loz = fminf(p0.z, p1.z), hiz = fmaxf(p0.z, p1.z);
  const float far = (fmaxf(fabsf(r.o.x - lox), fabsf(r.o.x - hix)) + fmaxf(fabsf(r.o.y - loy), fabsf(r.o.y - hiy))) +
                    fmaxf(fabsf(r.o.z - loz), fabsf(r.o.z - hiz));
  const float m = 1.0e-3f * (far + ((fabsf(r.o.x) + fabsf(r.o.y)) + fabsf(r.o.z)));
  if (!(dmin > 0.f && dmax >= kCullDirMin && dmax <= kCullDirMax && m - m == 0.f)) return false;  // (m: NaN / inf anywhere)
  const float ix = fminf(fmaxf(rcp_fast(r.d.x), -kCullInvMax), kCullInvMax), iy = fminf(fmaxf(rcp_fast(r.d.y), -kCullInvMax), kCullInvMax);
  const float iz = fminf(fmaxf(rcp_fast(r.d.z), -kCullInvMax), kCullInvMax);
  const float x0 = (lox - m - r.o.x) * ix, x1 = (hix + m - r.o.x) * ix, y0 = (loy - m - r.o.y) * iy, y1 = (hiy + m - r.o.y) * iy;
  const float z0 = (loz - m - r.o.z) * iz, z1 = (hiz + m - r.o.z) * iz;
  const float t_in = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
  const float t_out = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
  return t_in > t_out;
}

// constant_medium.hpp:28-78.  Draws one RNG number iff both boundary hits
// succeed and rec1.t < rec2.t after clipping.  `shortcut`: prove the common case -- the ray's line misses the boundary
// box altogether -- with one slab test instead of the twelve rectangle tests of the two boundary hits (off when culling
// is off, so that the tests compare with and without it).
PT_DEV bool medium_hit_t(const MediumRec& m, const Ray& r, float tmin, float tmax, Rng& rng, float& t_out, bool shortcut = true) {
  float t1, t2;
  if (shortcut && m.boundary_kind != PT_BOUNDARY_SPHERE && box_line_missed(r, vld(m.p0), vld(m.p1))) return false;
  if (m.boundary_kind == PT_BOUNDARY_SPHERE) {
    V3 center = vld(m.c0);
    if (m.moving) center = moving_center(center, vld(m.dv), fdiv(fsub(r.tm, m.time0), m.den));
    if (!sphere_hit_t(r, center, m.r2, -kInf, kInf, t1)) return false;
    if (!sphere_hit_t(r, center, m.r2, fadd(t1, 0.0001f), kInf, t2)) return false;
  } else {
    float a, b;
    const V3 p0 = vld(m.p0), p1 = vld(m.p1);
    if (box_hit_t(r, p0, p1, -kInf, kInf, t1, a, b) < 0) return false;
    if (box_hit_t(r, p0, p1, fadd(t1, 0.0001f), kInf, t2, a, b) < 0) return false;
  }
  if (t1 < tmin) t1 = tmin;
  if (t2 > tmax) t2 = tmax;
  if (t1 >= t2) return false;
  if (t1 < 0.f) t1 = 0.f;
  const float ray_length = vlength(r.d);
  const float distance_inside_boundary = fmul(fsub(t2, t1), ray_length);
  const float hit_distance = fmul(m.neg_inv_density, t_log(rng_float(rng)));
  if (hit_distance > distance_inside_boundary) return false;
  t_out = fadd(t1, fdiv(hit_distance, ray_length));
  return true;
}

// ---------------------------------------------------------------- the scan
// Merge the members' partial winners: afterwards every member of a team holds the team's winner.
PT_DEV void team_merge(const SceneDesc& sc, Best& best, int team_size) {
  for (int o = team_size >> 1; o > 0; o >>= 1) {
    const float ot = __shfl_xor_sync(0xffffffffu, best.t, o);
    const int oid = __shfl_xor_sync(0xffffffffu, best.id, o);
    if (oid >= 0) {
      if (best.id < 0 || ot < best.t)
        best.t = ot, best.id = oid;
      else if (ot == best.t && oid != best.id && key_of(sc, oid) > key_of(sc, best.id))
        best.id = oid;
    }
  }
}

// Two-phase sphere test.  Phase 1 is branch-free and only collects a bitmask of spheres whose
// discriminant is positive; phase 2 computes exact roots for the set bits.
// Centre of the sphere at slot `p` ({c, r*r} entry; a moving sphere's {c1 - c0} is 32 entries further,
// pt_packed.h) at the ray's time (sphere.hpp:51-56), exactly as the reference computes it.
template <bool kSmem, bool kMoving>
PT_DEV void sphere_center(const float4* __restrict__ p, float f, float& cx, float& cy, float& cz, float& r2_filter) {
  const float4 s = ld4<kSmem>(p);
  if constexpr (kMoving) {
    const float4 v = ld4<kSmem>(p + 2 * kSphereChunk);
    cx = fadd(s.x, fmul(f, v.x)), cy = fadd(s.y, fmul(f, v.y)), cz = fadd(s.z, fmul(f, v.z)), r2_filter = s.w;  // sphere.hpp:55
  } else {
    cx = s.x, cy = s.y, cz = s.z, r2_filter = s.w;
  }
}

// CONSERVATIVE MISS FILTER (the hot instruction sequence of the whole renderer).  The reference
// evaluates  disc = b*b - a*c,  b = dot(oc, d),  c = dot(oc, oc) - r*r  with 17 separately rounded
// operations and hits only if disc > 0 (sphere.hpp:68-74).  Parity needs that exact sequence ONLY
// for spheres that can be hit; for the others it is enough to PROVE disc <= 0.  The filter
// evaluates, with fused multiply-adds (11 operations),
//     test = b'^2 - a(1-k) * (|oc|^2 - r^2 (1+e)),   e = 2k / (1-k)
// which in exact arithmetic equals  disc + k * a * (|oc|^2 + r^2).  Either evaluation is within
// 13 u a (|oc|^2 + r^2) of the exact real value (u = 2^-24; error analysis in DESIGN.md), so with
// k = 4e-6 > 27 u the implication  (reference disc > 0)  =>  (test > 0)  always holds: the filter never
// drops a sphere the reference would hit.  Spheres that pass are re-evaluated with the exact
// sequence (sphere_roots_scan), so false positives only cost time.  The blob stores r^2 (1+e)
// (rounded up); the exact r*r comes from the side table.
constexpr float kFilterK = 4.0e-6f;
PT_DEV float filter_a(float a) { return fmul(a, 1.0f - kFilterK); }
// Returns the bits of -test: the SIGN BIT is set for every sphere the filter lets through (and, harmlessly,
// for -0 and some NaNs), so the per-lane candidate mask is collected with one funnel shift per sphere.
template <bool kSmem, bool kMoving>
PT_DEV uint32_t sphere_filter_bits(const float4* __restrict__ p, float f, const Ray& r, float a_filter) {
  float cx, cy, cz, r2f;
  sphere_center<kSmem, kMoving>(p, f, cx, cy, cz, r2f);
  const float ocx = fsub(r.o.x, cx), ocy = fsub(r.o.y, cy), ocz = fsub(r.o.z, cz);
  const float b = __fmaf_rn(ocx, r.d.x, __fmaf_rn(ocy, r.d.y, fmul(ocz, r.d.z)));
  const float c = __fmaf_rn(ocx, ocx, __fmaf_rn(ocy, ocy, __fmaf_rn(ocz, ocz, -r2f)));
  return __float_as_uint(__fmaf_rn(-b, b, fmul(a_filter, c)));
}
PT_DEV float exact_r2(const SphereAux* aux, int i) {
  const float radius = aux[i].radius;
  return fmul(radius, radius);  // sphere.hpp:71
}

// CHUNK CULLING.  Spheres come in chunks of 16 spatially close ones, each with a bounding box
// (pt_packed.h); a ray scans only the chunks whose box it crosses between t = 0 and the running
// closest hit.  The result is the reference's for ANY set of skipped chunks that cannot contain an
// accepted root (the winner rule is order independent), so what has to hold is: "the reference accepts
// a root of sphere i" => "the box test of i's chunk passes".  The boxes are grown on the host by a
// margin that covers the rounding of the reference's own root, of the centre and of this slab test
// for every origin with max |coordinate| <= cull_bound[set] (pt_pack.cpp, DESIGN.md "chunk
// culling"); the last set is infinite boxes and serves every other ray.  The slab test runs on
// 1/d clamped to +-2^60: for a component below 2^-60 the plane distances then come out within
// t / 2^60 of zero instead of exactly there, far inside the margin, as long as the direction's largest
// component lies in [2^-20, 2^20] -- rays outside that range are not culled at all.
struct CullRay {
  float ix, iy, iz;  // clamped 1 / d
  float qx, qy, qz;  // -o * (1 / d)
};
// Returns the box set of this ray.
PT_DEV int make_cull_ray(const SceneDesc& sc, const Ray& r, CullRay& c) {
  const float dmax = fmaxf(fmaxf(fabsf(r.d.x), fabsf(r.d.y)), fabsf(r.d.z));
  const float omax = fmaxf(fmaxf(fabsf(r.o.x), fabsf(r.o.y)), fabsf(r.o.z));
  int set = (kCullSets - 1) - ((omax <= sc.cull_bound[0]) + (omax <= sc.cull_bound[1]) + (omax <= sc.cull_bound[2]));
  if (!(dmax >= kCullDirMin && dmax <= kCullDirMax)) set = kCullSets - 1;
  c.ix = fminf(fmaxf(rcp_fast(r.d.x), -kCullInvMax), kCullInvMax);
  c.iy = fminf(fmaxf(rcp_fast(r.d.y), -kCullInvMax), kCullInvMax);
  c.iz = fminf(fmaxf(rcp_fast(r.d.z), -kCullInvMax), kCullInvMax);
  c.qx = -(r.o.x * c.ix), c.qy = -(r.o.y * c.iy), c.qz = -(r.o.z * c.iz);
  return set;
}
// Bits of (entry - exit) of the ray's interval inside the box, clipped to [0, tmax]: SIGN BIT set <=>
// the ray crosses the box there (fminf / fmaxf drop NaNs, which only ever widens the interval).
template <bool kSmem> PT_DEV uint32_t chunk_bits(const float4* __restrict__ box, const CullRay& c, float tmax) {
  const float4 lo = ld4<kSmem>(box), hi = ld4<kSmem>(box + 1);
  const float ax = __fmaf_rn(lo.x, c.ix, c.qx), bx = __fmaf_rn(hi.x, c.ix, c.qx);
  const float ay = __fmaf_rn(lo.y, c.iy, c.qy), by = __fmaf_rn(hi.y, c.iy, c.qy);
  const float az = __fmaf_rn(lo.z, c.iz, c.qz), bz = __fmaf_rn(hi.z, c.iz, c.qz);
  const float t_in = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
  const float t_out = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
  return __float_as_uint(t_in - t_out);
}

// Bitmask of the chunks cb .. cb + nb - 1 (nb <= 32) whose box the ray crosses before tmax: chunk cb + k at
// bit nb - 1 - k.  Same box for every lane: broadcast loads.
template <bool kSmem>
PT_DEV uint32_t chunk_hits(const float4* __restrict__ boxes, int cb, int nb, const CullRay& cr, float tmax) {
  uint32_t hits = 0;
#pragma unroll 2
  for (int k = 0; k < nb; ++k) hits = __funnelshift_l(chunk_bits<kSmem>(boxes + 2 * (cb + k), cr, tmax), hits, 1);
  return hits;
}

// Filter + exact roots of kOwn spheres of one chunk: slots p, p + kStep, ... of the doubled chunk, where
// p = chunk base + rot0 (pt_packed.h).  Lanes of a warp work on DIFFERENT chunks at the same time; with
// rot0 = (lane & 15) + const they read different shared-memory banks.
template <bool kSmem, bool kMoving, int kOwn, int kStep, typename Keys>
PT_DEV void scan_chunk(const Keys& sc, const float4* __restrict__ data, const SphereAux* aux, int chunk, int rot0,
                       const Ray& r, float a, float af, float f, int type, Best& best) {
  constexpr int kUnroll = kOwn < kScanUnroll ? kOwn : kScanUnroll;
  constexpr int kSlots = kMoving ? 4 * kSphereChunk : 2 * kSphereChunk;  // float4 per chunk
  const float4* base = data + chunk * kSlots + rot0;
  uint32_t mask = 0;  // own k-th sphere at bit kOwn - 1 - k
#pragma unroll 1
  for (int it = 0; it < kOwn; it += kUnroll) {
#pragma unroll
    for (int j = 0; j < kUnroll; ++j)
      mask = __funnelshift_l(sphere_filter_bits<kSmem, kMoving>(base + (it + j) * kStep, f, r, af), mask, 1);
  }
  while (mask) {
    const int mt = 31 - __clz((int)mask);
    mask &= ~(1u << mt);
    const int k = kOwn - 1 - mt;
    const int i = chunk * kSphereChunk + ((rot0 + k * kStep) & (kSphereChunk - 1));
    float cx, cy, cz, r2f;
    sphere_center<kSmem, kMoving>(base + k * kStep, f, cx, cy, cz, r2f);
    sphere_roots_scan(sc, best, r, a, cx, cy, cz, exact_r2(aux, i), make_id(type, i));
  }
}

// Scan the chunks [first_el / 16, end_el / 16) of one sphere group for one ray, kTeam lanes per ray: every
// lane first collects the bitmask of chunks the ray crosses, then pops its next chunk and filters its
// share of the 16 spheres (own k-th sphere = slot rot + k * kTeam, rot = lane & 15).
template <bool kSmem, bool kMoving, int kTeam, typename Keys>
PT_DEV void scan_sphere_chunks(const Keys& sc, const float4* __restrict__ data, const float4* __restrict__ boxes,
                               const SphereAux* aux, int first_el, int end_el, int rot, const Ray& r, float a,
                               float f, const CullRay& cr, bool act, int type, Best& best) {
  const float af = filter_a(a);
  const int c_end = end_el / kSphereChunk;
#pragma unroll 1
  for (int cb = first_el / kSphereChunk; cb < c_end; cb += 32) {
    const int nb = min(32, c_end - cb);
    float tmax = act ? best.t : -1.f;
    uint32_t hits = 0;  // chunk cb + k at bit nb - 1 - k
    if constexpr (kTeam == 1) {
      hits = chunk_hits<kSmem>(boxes, cb, nb, cr, tmax);
    } else {
      // The team shares the box tests: member m takes chunks m, m + kTeam, ... against the smallest of the
      // members' running closest hits, and the members' bits are OR-ed together (the whole warp is
      // converged here: the loop bounds depend on the group only).
#pragma unroll
      for (int o = kTeam >> 1; o > 0; o >>= 1) tmax = fminf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
#pragma unroll 1
      for (int k = rot & (kTeam - 1); k < nb; k += kTeam)
        hits |= (chunk_bits<kSmem>(boxes + 2 * (cb + k), cr, tmax) >> 31) << (nb - 1 - k);
#pragma unroll
      for (int o = kTeam >> 1; o > 0; o >>= 1) hits |= __shfl_xor_sync(0xffffffffu, hits, o);
    }
#pragma unroll 1
    while (hits) {
      const int top = 31 - __clz((int)hits);
      hits &= ~(1u << top);
      scan_chunk<kSmem, kMoving, kSphereChunk / kTeam, kTeam>(sc, data, aux, cb + (nb - 1 - top), rot, r, a, af, f, type, best);
    }
  }
}

// The team variants (kTeam lanes share a ray, each with its own partial winner): out of line and by
// value, so that they do not sit between the hot loops in the instruction stream.
template <bool kSmem, bool kMoving>
__device__ __noinline__ Best scan_spheres_team(KeyTable sc, const float4* __restrict__ data,
                                               const float4* __restrict__ boxes, const SphereAux* aux, int first_el,
                                               int end_el, int team_size, Ray r, float a, float f, CullRay cr, bool act,
                                               int type, Best best) {
  const int rot = (int)(threadIdx.x & (kSphereChunk - 1));
  switch (team_size) {
    case 2: scan_sphere_chunks<kSmem, kMoving, 2>(sc, data, boxes, aux, first_el, end_el, rot, r, a, f, cr, act, type, best); break;
    case 4: scan_sphere_chunks<kSmem, kMoving, 4>(sc, data, boxes, aux, first_el, end_el, rot, r, a, f, cr, act, type, best); break;
    case 8: scan_sphere_chunks<kSmem, kMoving, 8>(sc, data, boxes, aux, first_el, end_el, rot, r, a, f, cr, act, type, best); break;
    default: scan_sphere_chunks<kSmem, kMoving, 16>(sc, data, boxes, aux, first_el, end_el, rot, r, a, f, cr, act, type, best); break;
  }
  return best;
}

template <bool kSmem, bool kMoving>
PT_DEV void scan_spheres(const SceneDesc& sc, const float4* __restrict__ data, const float4* __restrict__ boxes,
                         const SphereAux* aux, int first_el, int end_el, int team_size, const Ray& r, float a, float f,
                         const CullRay& cr, bool act, int type, Best& best) {
  if (team_size == 1)
    scan_sphere_chunks<kSmem, kMoving, 1>(sc, data, boxes, aux, first_el, end_el, (int)(threadIdx.x & (kSphereChunk - 1)),
                                          r, a, f, cr, act, type, best);
  else
    best = scan_spheres_team<kSmem, kMoving>(key_table(sc), data, boxes, aux, first_el, end_el, team_size, r, a, f, cr,
                                             act, type, best);
}

// Rectangles, triangles or boxes: elements first, first + step, ... below `end` (running closest as the upper
// bound, like the reference's loop).
template <bool kSmem, typename Keys>
PT_DEV void scan_flat_range(const Keys& sc, const float4* __restrict__ data, int type, int first, int end, int step, const Ray& r,
                            Best& best) {
  if (type == G_RECT) {
    for (int i = first; i < end; i += step) {
      const float4 q0 = ld4<kSmem>(data + 2 * i);
      const float4 q1 = ld4<kSmem>(data + 2 * i + 1);
      float t, ra, rb;
      if (rect_hit_t(r, __float_as_int(q1.y), q0.x, q0.y, q0.z, q0.w, q1.x, kTMin, best.t, t, ra, rb))
        consider_le(sc, best, t, make_id(G_RECT, i));
    }
  } else if (type == G_TRIANGLE) {
    for (int i = first; i < end; i += step) {
      PT_STAT(triangle_tests);
      const float4 v0 = ld4<kSmem>(data + 3 * i);
      const float4 e1 = ld4<kSmem>(data + 3 * i + 1);
      const float4 e2 = ld4<kSmem>(data + 3 * i + 2);
      float t;
      if (triangle_hit_t(r, v3(v0.x, v0.y, v0.z), v3(e1.x, e1.y, e1.z), v3(e2.x, e2.y, e2.z), kTMin, best.t, t))
        consider_le(sc, best, t, make_id(G_TRIANGLE, i));
    }
  } else if (type == G_BOX) {
    for (int i = first; i < end; i += step) {
      const float4 p0 = ld4<kSmem>(data + 2 * i);
      const float4 p1 = ld4<kSmem>(data + 2 * i + 1);
      float t, ra, rb;
      if (box_hit_t(r, v3(p0.x, p0.y, p0.z), v3(p1.x, p1.y, p1.z), kTMin, best.t, t, ra, rb) >= 0)
        consider_le(sc, best, t, make_id(G_BOX, i));
    }
  }
}
PT_DEV bool has_tree(const SceneDesc& sc, const Group& g) { return g.tree >= 0 && sc.flat_cull != 0u; }
PT_DEV const float4* flat_data(const SceneView& sv, int type) {
  return type == G_RECT ? sv.rect() : type == G_TRIANGLE ? sv.triangle() : sv.box();
}
template <bool kSmem>
PT_DEV void scan_flat_group(const SceneDesc& sc, const SceneView& sv, const Group& g, const Ray& r, int first, int step,
                            Best& best) {
  scan_flat_range<kSmem>(sc, flat_data(sv, g.type), g.type, first, g.begin + g.count, step, r, best);
}

// FLAT CULLING.  A flat group with a tree (pt_packed.h) is scanned leaf by leaf, and only the leaves whose
// box the ray crosses.  What has to hold, as for the sphere chunks: "the reference accepts element i at
// parameter t" => "the box test of every node above i passes".  DESIGN.md ("flat culling") shows that the
// point o + t d of an accepted hit lies within
//     M = kFlatRel * far + kFlatAbs * (max |o_k| + extent),   far = distance from o to the box's far corner,
// of the element's bounding box -- for a rectangle or a box side because the reference itself checks the hit
// point against the bounds (rectangle.hpp:38-41), for a triangle as long as the Moller-Trumbore determinant
// satisfies |a| >= kGrazeTau |d| |e1| |e2| (the residual of the float solution in the exact equations is
// 24 u |s| |d| |e1| |e2| / |a|, u = 2^-24).  Every node is therefore tested with its box grown by M (M only
// grows towards the root, so a crossed leaf has crossed ancestors).  Rays whose direction or origin is out
// of the range the slab test is proven for are not culled at all (FlatRay::ok).
// GRAZING INDEX.  Below that determinant bound the reference's u, v and t are rounding noise and it may
// "hit" a triangle the ray is nowhere near.  Such pairs satisfy |d . g| < (kGrazeTau + 7 u) |d| with g =
// cross(e1, e2) / (|e1| |e2|): they are found in a second tree over the g vectors and tested exactly, box
// or no box.  (Testing a triangle twice is harmless: the winner rule is idempotent.)
#ifndef PT_FLAT_REL  // (overridden only by the sensitivity experiments of tests/host: do mismatches appear when they should?)
#define PT_FLAT_REL 3.2e-3f
#define PT_FLAT_ABS 4.0e-6f
#define PT_GRAZE_TAU 4.8828125e-4f
#endif
constexpr float kFlatRel = PT_FLAT_REL;
constexpr float kFlatAbs = PT_FLAT_ABS;
constexpr float kGrazeTau = PT_GRAZE_TAU;                     // 2^-11
constexpr float kGrazeTauDev = kGrazeTau + 64.f * 5.9604645e-8f;  // + the rounding of g and of the interval product
struct FlatRay {
  CullRay c;
  float eabs;  // kFlatAbs * (max |o_k| + extent)
  float taud;  // grazing threshold: kGrazeTauDev * |d|, rounded up
  bool ok;     // false: this ray visits every leaf
};
PT_DEV float sqrt_fast(float x) {
#ifdef __CUDACC__
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
PT_DEV float graze_threshold(const Ray& r) {
  return kGrazeTauDev * 1.0001f * sqrt_fast(__fmaf_rn(r.d.x, r.d.x, __fmaf_rn(r.d.y, r.d.y, r.d.z * r.d.z)));
}
PT_DEV FlatRay make_flat_ray(float flat_extent, const Ray& r) {
  FlatRay f;
  const float dmax = fmaxf(fmaxf(fabsf(r.d.x), fabsf(r.d.y)), fabsf(r.d.z));
  const float omax = fmaxf(fmaxf(fabsf(r.o.x), fabsf(r.o.y)), fabsf(r.o.z));
  // (a ray with an exactly zero direction component lies IN the planes k = o_k: a rectangle or box side there has
  // t = 0 / 0 = NaN, which the reference accepts wherever the rectangle is, rectangle.hpp:35-41 -- such rays, NaN
  // and far-out origins, and directions the clamped reciprocal does not cover are never culled)
  f.ok = dmax >= kCullDirMin && dmax <= kCullDirMax && omax <= 1.0e15f && r.o.x == r.o.x && r.o.y == r.o.y && r.o.z == r.o.z &&
         r.d.x == r.d.x && r.d.y == r.d.y && r.d.z == r.d.z && r.d.x != 0.f && r.d.y != 0.f && r.d.z != 0.f;
  f.c.ix = fminf(fmaxf(rcp_fast(r.d.x), -kCullInvMax), kCullInvMax);
  f.c.iy = fminf(fmaxf(rcp_fast(r.d.y), -kCullInvMax), kCullInvMax);
  f.c.iz = fminf(fmaxf(rcp_fast(r.d.z), -kCullInvMax), kCullInvMax);
  f.c.qx = -(r.o.x * f.c.ix), f.c.qy = -(r.o.y * f.c.iy), f.c.qz = -(r.o.z * f.c.iz);
  f.eabs = kFlatAbs * (omax + flat_extent);
  f.taud = graze_threshold(r);
  return f;
}
// SIGN BIT set <=> the ray crosses the node's box, grown by M, between t = 0 and tmax.
template <bool kSmem> PT_DEV uint32_t flat_node_bits(const float4* __restrict__ box, const FlatRay& f, const Ray& r, float tmax) {
  PT_STAT(flat_nodes);
  const float4 lo = ld4<kSmem>(box), hi = ld4<kSmem>(box + 1);
  const float fx = fmaxf(fabsf(r.o.x - lo.x), fabsf(r.o.x - hi.x));
  const float fy = fmaxf(fabsf(r.o.y - lo.y), fabsf(r.o.y - hi.y));
  const float fz = fmaxf(fabsf(r.o.z - lo.z), fabsf(r.o.z - hi.z));
  const float far = sqrt_fast(__fmaf_rn(fx, fx, __fmaf_rn(fy, fy, fz * fz)));
  const float m = __fmaf_rn(kFlatRel, far, f.eabs);
  const float ax = __fmaf_rn(lo.x - m, f.c.ix, f.c.qx), bx = __fmaf_rn(hi.x + m, f.c.ix, f.c.qx);
  const float ay = __fmaf_rn(lo.y - m, f.c.iy, f.c.qy), by = __fmaf_rn(hi.y + m, f.c.iy, f.c.qy);
  const float az = __fmaf_rn(lo.z - m, f.c.iz, f.c.qz), bz = __fmaf_rn(hi.z + m, f.c.iz, f.c.qz);
  const float t_in = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
  const float t_out = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
  return __float_as_uint(t_in - t_out) | (f.ok ? 0u : 0x80000000u);
}
// SIGN BIT set <=> some g inside the node's box may have |d . g| <= taud (interval product).
template <bool kSmem> PT_DEV uint32_t graze_node_bits(const float4* __restrict__ box, const V3& d, float taud) {
  PT_STAT(graze_nodes);
  const float4 lo = ld4<kSmem>(box), hi = ld4<kSmem>(box + 1);
  const float x0 = d.x * lo.x, x1 = d.x * hi.x, y0 = d.y * lo.y, y1 = d.y * hi.y, z0 = d.z * lo.z, z1 = d.z * hi.z;
  const float mn = (fminf(x0, x1) + fminf(y0, y1)) + fminf(z0, z1);
  const float mx = (fmaxf(x0, x1) + fmaxf(y0, y1)) + fmaxf(z0, z1);
  return (mn <= taud && mx >= -taud) ? 0x80000000u : 0u;
}

// Level l of a tree by name, not by index (a dynamically indexed struct would live in local memory).
PT_DEV int tree_off(const Tree& t, int l) { return l == 0 ? t.off[0] : l == 1 ? t.off[1] : t.off[2]; }
PT_DEV int tree_n(const Tree& t, int l) { return l == 0 ? t.n[0] : l == 1 ? t.n[1] : t.n[2]; }
// Visit the leaves under the nodes [first, first + count) of level `top` whose ancestors below `top` (and whose own
// box) all pass `test` (sign bit of its result).  The whole tree: top = t.levels - 1, first = 0, count = n[top].
template <typename Test, typename Leaf>
PT_DEV void tree_walk(const Tree& t, const float4* __restrict__ nodes, int top, int first, int count, Test test, Leaf leaf) {
  // A = the nodes of level `top`, B = their children, C = theirs
  const int off_a = tree_off(t, top);
  const int off_b = top == 2 ? t.off[1] : t.off[0], n_b = top == 2 ? t.n[1] : t.n[0];
  const float4* b2 = nodes + off_a;
  const int n2 = first + count;
#pragma unroll 1
  for (int c2 = first; c2 < n2; c2 += 32) {
    const int k2 = min(32, n2 - c2);
    uint32_t m2 = 0;
#pragma unroll 1
    for (int k = 0; k < k2; ++k) m2 = __funnelshift_l(test(b2 + 2 * (c2 + k)), m2, 1);
#pragma unroll 1
    while (m2) {
      const int p2 = 31 - __clz((int)m2);
      m2 &= ~(1u << p2);
      const int i2 = c2 + (k2 - 1 - p2);
      if (top == 0) {
        leaf(i2);
        continue;
      }
      const float4* b1 = nodes + off_b;
      const int c1 = i2 * kTreeFan, k1 = min(kTreeFan, n_b - c1);
      uint32_t m1 = 0;
#pragma unroll 1
      for (int k = 0; k < k1; ++k) m1 = __funnelshift_l(test(b1 + 2 * (c1 + k)), m1, 1);
#pragma unroll 1
      while (m1) {
        const int p1 = 31 - __clz((int)m1);
        m1 &= ~(1u << p1);
        const int i1 = c1 + (k1 - 1 - p1);
        if (top == 1) {
          leaf(i1);
          continue;
        }
        const float4* b0 = nodes + t.off[0];
        const int c0 = i1 * kTreeFan, k0 = min(kTreeFan, t.n[0] - c0);
        uint32_t m0 = 0;
#pragma unroll 1
        for (int k = 0; k < k0; ++k) m0 = __funnelshift_l(test(b0 + 2 * (c0 + k)), m0, 1);
#pragma unroll 1
        while (m0) {
          const int p0 = 31 - __clz((int)m0);
          m0 &= ~(1u << p0);
          leaf(c0 + (k0 - 1 - p0));
        }
      }
    }
  }
}

// One leaf of the grazing index: its triangles, wherever they lie.
template <bool kSmem, typename Keys>
PT_DEV void scan_graze_leaf(const Keys& sc, const float4* __restrict__ triangles, const float4* __restrict__ entries, int first, int step,
                            const Ray& r, float taud, Best& best) {
  for (int j = first; j < kFlatChunk; j += step) {
    const float4 e = ld4<kSmem>(entries + j);
    const int i = __float_as_int(e.w);
    PT_STAT(graze_tests);
    if (i >= 0 && fabsf(__fmaf_rn(r.d.x, e.x, __fmaf_rn(r.d.y, e.y, r.d.z * e.z))) <= taud)
      scan_flat_range<kSmem>(sc, triangles, G_TRIANGLE, i, i + 1, 1, r, best);
  }
}

// A flat group with a tree for one ray, in place (lane kernel, sequential fallback, LATE): member `first` of a
// team of `step` takes every step-th element of each visited leaf.  Out of line and by value (one copy per
// kernel, nothing forced into local memory).
struct FlatTrees {
  const float4* data;  // the group's kind's array
  const Tree* trees;
  const float4* nodes;
  const float4* ids;
  float extent;
};
PT_DEV FlatTrees flat_trees(const SceneDesc& sc, const SceneView& sv, int type) {
  return FlatTrees { flat_data(sv, type), sv.trees(), sv.nodes(), sv.tree_ids(), sc.flat_extent };
}
template <bool kSmem>
__device__ __noinline__ Best scan_flat_tree(KeyTable sc, FlatTrees ft, Group g, Ray r, int first, int step, Best best) {
  const FlatRay fr = make_flat_ray(ft.extent, r);
  const Tree t = ft.trees[g.tree];
  tree_walk(
      t, ft.nodes, t.levels - 1, 0, tree_n(t, t.levels - 1), [&](const float4* box) { return flat_node_bits<kSmem>(box, fr, r, best.t); },
      [&](int leaf) {
        const int b = g.begin + leaf * kFlatChunk;
        scan_flat_range<kSmem>(sc, ft.data, g.type, b + first, min(b + kFlatChunk, g.begin + g.count), step, r, best);
      });
  if (g.gtree >= 0 && fr.ok) {
    const Tree gt = ft.trees[g.gtree];
    tree_walk(
        gt, ft.nodes, gt.levels - 1, 0, tree_n(gt, gt.levels - 1), [&](const float4* box) { return graze_node_bits<kSmem>(box, r.d, fr.taud); },
        [&](int leaf) { scan_graze_leaf<kSmem>(sc, ft.data, ft.ids + gt.leaf_ids + leaf * kFlatChunk, first, step, r, fr.taud, best); });
  }
  return best;
}

// THE SCAN IN VECTOR ORDER.  The (minimum t, maximum key) rule is the sequential scan's result only while every t
// is a number.  A rectangle or box side whose plane contains the ray (d_k == 0 exactly and o_k == k) has t = 0 / 0 =
// NaN, which the reference ACCEPTS (rectangle.hpp:35-41 reject with `t < min || t > max`) and which then poisons its
// running closest hit: every later sphere is rejected (`temp < NaN`), every later flat object accepted.  No order-
// independent rule reproduces that, so rays that can meet a NaN -- an exactly zero direction component, or anything
// not finite -- are scanned the reference's way: object by object in vector order against the running closest hit.
// So are rays whose direction is outside [2^-20, 2^20]: the error bounds behind the miss filter and the box tests
// assume no underflow or overflow in d . d (found by tests/host/scan_check.cpp: a direction of length 1e-23).
// Out of line (it is rare: pixel (0, 0), whose generator returns zeros forever, and hand-made cameras).
// Does the ray with d_c == 0 start ON a plane that carries a rectangle / box side (SceneDesc::off_planes)?  Out of line: once
// in millions of rays.
static __device__ __noinline__ bool on_a_plane(const SceneDesc* sc, const unsigned char* blob_base, int c, float o_c) {
  const float* v = reinterpret_cast<const float*>(blob_base + sc->off_planes);
  int first = 0;
  for (int k = 0; k < c; ++k) first += (int)sc->n_planes[k];
  const int n = (int)sc->n_planes[c];
  for (int i = 0; i < n; ++i)
    if (v[first + i] == o_c) return true;
  return false;
}
PT_DEV bool needs_in_order(const SceneDesc& sc, const unsigned char* blob_base, const Ray& r) {
#ifdef PT_NO_INORDER  // (experiments only)
  return false;
#endif
  const float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
  const float dmax = fmaxf(fmaxf(ax, ay), az), dmin = fminf(fminf(ax, ay), az);
  const float all = ((r.o.x + r.o.y) + r.o.z) + ((r.d.x + r.d.y) + r.d.z);  // (fminf / fmaxf drop a NaN: catch it here; inf - inf is NaN)
  if (dmin > 0.f && dmax >= kCullDirMin && dmax <= kCullDirMax && all - all == 0.f) return false;  // (all but one ray in millions)
  if (!(dmax >= kCullDirMin && dmax <= kCullDirMax && all - all == 0.f)) return true;
  // a zero component: the only NaN a finite ray can meet is t = (k - o_c) / d_c = 0 / 0 on a plane it starts on; off every
  // plane t is +-inf, the hit point's other coordinates are +-inf and the rectangle's bounds reject it -- unless a SECOND
  // component is zero too (inf x 0 = NaN passes the bounds, rectangle.hpp:38-41)
  const int zeros = (ax == 0.f) + (ay == 0.f) + (az == 0.f);
  if (zeros != 1) return true;
  return ax == 0.f ? on_a_plane(&sc, blob_base, 0, r.o.x) : ay == 0.f ? on_a_plane(&sc, blob_base, 1, r.o.y) : on_a_plane(&sc, blob_base, 2, r.o.z);
}
template <bool kSmem>
__device__ __noinline__ Best closest_hit_in_order(const SceneDesc* scp, const float4* sphere, const float4* moving, const float4* rect,
                                                  const float4* triangle, const float4* box, Ray r, Rng* rng) {
  const SceneDesc& sc = *scp;
  Best best { kInf, -1 };
  float closest = kInf;  // render.hpp:35
  const int n = (int)sc.n_objects;
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const int id = sc.object_id[i];
    if (id < 0) continue;
    const int idx = id & (int)kIdMask;
    float t, ra, rb;
    bool hit = false;
    switch (id >> kIdShift) {
      case G_SPHERE: {
        const float4 s = ld4<kSmem>(sphere + sphere_slot(idx));
        hit = sphere_hit_t(r, v3(s.x, s.y, s.z), exact_r2(sc.sphere_aux, idx), kTMin, closest, t);
        break;
      }
      case G_MOVING_SPHERE: {
        const float4 s = ld4<kSmem>(moving + moving_slot(idx)), v = ld4<kSmem>(moving + moving_slot(idx) + 2 * kSphereChunk);
        const SphereAux* aux = sc.moving_aux + idx;
        const V3 center = moving_center(v3(s.x, s.y, s.z), v3(v.x, v.y, v.z), fdiv(fsub(r.tm, aux->time0), aux->den));
        hit = sphere_hit_t(r, center, exact_r2(sc.moving_aux, idx), kTMin, closest, t);
        break;
      }
      case G_RECT: {
        const float4 q0 = ld4<kSmem>(rect + 2 * idx), q1 = ld4<kSmem>(rect + 2 * idx + 1);
        hit = rect_hit_t(r, __float_as_int(q1.y), q0.x, q0.y, q0.z, q0.w, q1.x, kTMin, closest, t, ra, rb);
        break;
      }
      case G_TRIANGLE: {
        const float4 v0 = ld4<kSmem>(triangle + 3 * idx), e1 = ld4<kSmem>(triangle + 3 * idx + 1), e2 = ld4<kSmem>(triangle + 3 * idx + 2);
        hit = triangle_hit_t(r, v3(v0.x, v0.y, v0.z), v3(e1.x, e1.y, e1.z), v3(e2.x, e2.y, e2.z), kTMin, closest, t);
        break;
      }
      case G_BOX: {
        const float4 p0 = ld4<kSmem>(box + 2 * idx), p1 = ld4<kSmem>(box + 2 * idx + 1);
        hit = box_hit_t(r, v3(p0.x, p0.y, p0.z), v3(p1.x, p1.y, p1.z), kTMin, closest, t, ra, rb) >= 0;
        break;
      }
      default:
        hit = medium_hit_t(sc.media[idx], r, kTMin, closest, *rng, t, false);
        break;
    }
    if (hit) closest = t, best.t = t, best.id = id;  // render.hpp:44-47
  }
  return best;
}
template <bool kSmem> PT_DEV Best closest_hit_in_order(const SceneDesc& sc, const SceneView& sv, const Ray& r, Rng& rng) {
  return closest_hit_in_order<kSmem>(&sc, sv.sphere(), sv.moving(), sv.rect(), sv.triangle(), sv.box(), r, &rng);
}

// render.hpp:30-51 for one ray per TEAM: `member` in [0, team_size) takes every team_size-th object
// of each group.  `act`: this lane's team carries a real ray (the others ride along so that the
// warp stays converged on the shared loads and the shuffles).
template <bool kSmem, bool kTrees = true>
PT_DEV Best closest_hit(const SceneDesc& sc, const SceneView& sv, const Ray& r, Rng& rng, bool act, int member,
                        int team_size) {
  // a ray that can meet a NaN is scanned in vector order (above) by every member of its team, AFTER the team has
  // ridden along here without a ray: the other teams of the warp need it for their full-warp shuffles
  const bool in_order = act && needs_in_order(sc, sv.base, r);
  act = act && !in_order;
  Best best { kInf, -1 };
  const float a = vdot(r.d, r.d);  // sphere.hpp:69, loop invariant
  CullRay cr;
  const int cull_set = make_cull_ray(sc, r, cr);
  const float4* sphere_boxes = sv.sphere_box() + cull_set * 2 * (int)sc.n_sphere_chunks;
  const float4* moving_boxes = sv.moving_box() + cull_set * 2 * (int)sc.n_moving_chunks;
  const int n_groups = (int)sc.n_groups;
  for (int gi = 0; gi < n_groups; ++gi) {
    const Group g = sv.groups()[gi];
    const int end = g.begin + g.count;
    switch (g.type) {
      case G_SPHERE:
        scan_spheres<kSmem, false>(sc, sv.sphere(), sphere_boxes, sc.sphere_aux, g.begin, end, team_size, r, a, 0.f, cr, act,
                                   G_SPHERE, best);
        break;
      case G_MOVING_SPHERE:
        scan_spheres<kSmem, true>(sc, sv.moving(), moving_boxes, sc.moving_aux, g.begin, end, team_size, r, a,
                                  fdiv(fsub(r.tm, g.time0), g.den), cr, act, G_MOVING_SPHERE, best);
        break;
      case G_MEDIUM: {
        // G_MEDIUM sees the running closest hit of EVERY lower-index object (merge first) and commits
        // unconditionally; every member replays the RNG draw on its replica of the generator.
        if (team_size > 1) team_merge(sc, best, team_size);
        if (act) {
          float t;
          if (medium_hit_t(sc.media[g.begin], r, kTMin, best.t, rng, t, sc.flat_cull != 0u)) best.t = t, best.id = make_id(G_MEDIUM, g.begin);
        }
        break;
      }
      default:
        if (act) {
          if (kTrees && has_tree(sc, g))
            best = scan_flat_tree<kSmem>(key_table(sc), flat_trees(sc, sv, g.type), g, r, member, team_size, best);
          else
            scan_flat_group<kSmem>(sc, sv, g, r, g.begin + member, team_size, best);
        }
        break;
    }
  }
  if (team_size > 1) team_merge(sc, best, team_size);
  if (in_order) best = closest_hit_in_order<kSmem>(sc, sv, r, rng);
  return best;
}

// ---------------------------------------------------------------- one ray per WARP (the deep-pixel lane of pt_wave.cu)
// All 32 lanes hold the same ray.  Sphere groups: lane l tests the box of chunk cb + l, the crossed chunks come back
// as a ballot and are scanned two at a time -- lanes 0-15 one sphere each of the first, lanes 16-31 of the second --
// with the usual two-phase test.  Every lane keeps a partial winner (its own running closest hit is a valid upper
// bound for the boxes it tests); team_merge() joins them under the (t, key) rule.
template <bool kSmem, bool kMoving>
PT_DEV void scan_sphere_chunks_warp(const SceneDesc& sc, const float4* __restrict__ data, const float4* __restrict__ boxes,
                                    const SphereAux* aux, int first_el, int end_el, int lane, const Ray& r, float a, float f,
                                    const CullRay& cr, int type, Best& best) {
  constexpr int kSlots = kMoving ? 4 * kSphereChunk : 2 * kSphereChunk;  // float4 per chunk
  const float af = filter_a(a);
  const int c_end = end_el / kSphereChunk;
  const int half = lane >> 4, j = lane & (kSphereChunk - 1);
#pragma unroll 1
  for (int cb = first_el / kSphereChunk; cb < c_end; cb += 32) {
    const int c = cb + lane;
    unsigned hits = __ballot_sync(0xffffffffu, c < c_end && (int)chunk_bits<kSmem>(boxes + 2 * c, cr, best.t) < 0);
#pragma unroll 1
    while (hits) {
      const int c0 = __ffs((int)hits) - 1;
      hits &= hits - 1u;
      int mine = cb + c0;
      if (hits) {
        const int c1 = __ffs((int)hits) - 1;
        hits &= hits - 1u;
        if (half) mine = cb + c1;
      } else if (half) {
        mine = -1;
      }
      if (mine >= 0) {
        const float4* p = data + mine * kSlots + j;
        if ((int)sphere_filter_bits<kSmem, kMoving>(p, f, r, af) < 0) {
          float cx, cy, cz, r2f;
          sphere_center<kSmem, kMoving>(p, f, cx, cy, cz, r2f);
          const int i = mine * kSphereChunk + j;
          sphere_roots_scan(sc, best, r, a, cx, cy, cz, exact_r2(aux, i), make_id(type, i));
        }
      }
    }
  }
}

// render.hpp:30-51 for one ray held by all 32 lanes of a warp: closest_hit() with a team of 32, the sphere groups
// scanned warp-wide (above), flat objects one per lane, media replayed by every lane on its replica of the generator.
template <bool kSmem, bool kTrees>
PT_DEV Best closest_hit_warp(const SceneDesc& sc, const SceneView& sv, const Ray& r, Rng& rng, int lane) {
  Best best { kInf, -1 };
  const float a = vdot(r.d, r.d);  // sphere.hpp:69, loop invariant
  CullRay cr;
  const int cull_set = make_cull_ray(sc, r, cr);
  const int n_groups = (int)sc.n_groups;
  bool dirty = false;  // partial winners not merged yet
#pragma unroll 1
  for (int gi = 0; gi < n_groups; ++gi) {
    const Group g = sv.groups()[gi];
    const int end = g.begin + g.count;
    if (g.type == G_SPHERE) {
      scan_sphere_chunks_warp<kSmem, false>(sc, sv.sphere(), sv.sphere_box() + cull_set * 2 * (int)sc.n_sphere_chunks, sc.sphere_aux,
                                            g.begin, end, lane, r, a, 0.f, cr, G_SPHERE, best);
      dirty = true;
    } else if (g.type == G_MOVING_SPHERE) {
      scan_sphere_chunks_warp<kSmem, true>(sc, sv.moving(), sv.moving_box() + cull_set * 2 * (int)sc.n_moving_chunks, sc.moving_aux,
                                           g.begin, end, lane, r, a, fdiv(fsub(r.tm, g.time0), g.den), cr, G_MOVING_SPHERE, best);
      dirty = true;
    } else if (g.type == G_MEDIUM) {
      // the medium sees the running closest hit of EVERY lower-index object (merge first) and commits unconditionally
      if (dirty) team_merge(sc, best, 32), dirty = false;
      float t;
      if (medium_hit_t(sc.media[g.begin], r, kTMin, best.t, rng, t)) best.t = t, best.id = make_id(G_MEDIUM, g.begin);
    } else {
      if (kTrees && has_tree(sc, g))
        best = scan_flat_tree<kSmem>(key_table(sc), flat_trees(sc, sv, g.type), g, r, lane, 32, best);
      else
        scan_flat_group<kSmem>(sc, sv, g, r, g.begin + lane, 32, best);
      dirty = true;
    }
  }
  if (dirty) team_merge(sc, best, 32);
  return best;
}

}  // namespace
}  // namespace ptb
#endif
