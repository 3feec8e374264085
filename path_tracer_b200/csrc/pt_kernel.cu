// pt_kernel.cu -- the per-pixel path-tracing loop of triSYCL/path_tracer
// (reference include/render.hpp:25-106 and everything it calls) as persistent sm_100a kernels.
//
// Execution model (B200-first, not a translation of the SYCL kernel):
//   * A pixel's `spp` samples are inherently serial -- its xorshift32 stream is consumed in a
//     data-dependent way (render.hpp:95-101) -- so the unit of work is a PIXEL, pulled from a global
//     atomic queue; a finished pixel is replaced at once ("path regeneration").  Parallelism is
//     pixels x objects.
//   * The scene's scan blob (pt_packed.h) and its side tables are staged into shared memory with
//     cp.async.bulk (TMA) once per CTA.  Spheres come in k-d ordered CHUNKS of 16 behind conservative
//     bounding boxes; a ray only looks at the chunks whose box it crosses ("chunk culling" below: a proof,
//     the result is bit-identical with and without it).  Sphere tests are two-phase: a branch-free FMA
//     filter that only records a candidate bitmask, and the exact roots (sqrt, IEEE division, range and
//     tie rules) over the few set bits.  Only `t` and the object id are tracked; the hit_record is
//     rebuilt once, for the winner.
//   * The winner of a scan is (minimum t, then maximum key), which reproduces the sequential scan's tie
//     behaviour for ANY visiting order (pt_packed.h): that is what allows skipping chunks, splitting a
//     scan over lanes or work items, and merging with shuffles or one 64-bit atomicMin.
//   * render_wave_kernel (default): one 896-thread CTA per SM, the path state of 896 pixels in a
//     structure-of-arrays pool in shared memory, phases BOXES / SPHERES / LATE / SHADE separated by
//     __syncthreads(); the closest-hit scan runs as (ray, chunk) work items spread over the whole CTA;
//     shading runs in warps of one material kind.  Deep pixels (paths bouncing dozens of times inside
//     glass hold ten times the average work and would sit on a serial critical path) are traced by
//     EXPRESS CTAs in short rounds, from the head of a longest-processing-time-first pixel order and
//     from a global hand-off queue.
//   * render_kernel (lane kernel): a warp holds k = 32 / T pixels, each owned by a TEAM of T lanes that
//     carry the pixel's path state REPLICATED; the scan is split inside a team and merged with shuffles,
//     bit-identical for every T.  The simpler scheduler, and the scan of the wavefront kernel's
//     sequential fallback.
// All arithmetic follows the operation order of the reference; see pt_device.cuh for the
// numerics contract.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pt_abi.h"
#include "pt_device.cuh"
#include "pt_kernel.h"
#include "pt_packed.h"

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

PT_DEV unsigned int ld_volatile_u32(const unsigned int* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }

PT_DEV unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- staging
PT_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

PT_DEV void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PT_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
PT_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
PT_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------- scene view
struct SceneView {
  const Group* groups;
  const float4* sphere;
  const float4* moving;
  const float4* rect;
  const float4* triangle;
  const float4* box;
  const float4* sphere_box;  // chunk boxes, set 0 first
  const float4* moving_box;
};

template <bool kSmem> PT_DEV float4 ld4(const float4* p) {
  if constexpr (kSmem)
    return *p;
  else
    return __ldg(p);
}

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

// constant_medium.hpp:28-78.  Draws one RNG number iff both boundary hits
// succeed and rec1.t < rec2.t after clipping.
PT_DEV bool medium_hit_t(const MediumRec& m, const Ray& r, float tmin, float tmax, Rng& rng, float& t_out) {
  float t1, t2;
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
constexpr float kCullInvMax = 1.152921504606847e18f;  // 2^60
constexpr float kCullDirMin = 9.5367431640625e-7f;    // 2^-20
constexpr float kCullDirMax = 1048576.f;              // 2^20
PT_DEV float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
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

// Rectangles, triangles and boxes of one group: elements first, first + step, ... (running closest as the
// upper bound, like the reference's loop).
template <bool kSmem>
PT_DEV void scan_flat_group(const SceneDesc& sc, const SceneView& sv, const Group& g, const Ray& r, int first, int step,
                            Best& best) {
  const int end = g.begin + g.count;
  if (g.type == G_RECT) {
    for (int i = first; i < end; i += step) {
      const float4 q0 = ld4<kSmem>(sv.rect + 2 * i);
      const float4 q1 = ld4<kSmem>(sv.rect + 2 * i + 1);
      float t, ra, rb;
      if (rect_hit_t(r, __float_as_int(q1.y), q0.x, q0.y, q0.z, q0.w, q1.x, kTMin, best.t, t, ra, rb))
        consider_le(sc, best, t, make_id(G_RECT, i));
    }
  } else if (g.type == G_TRIANGLE) {
    for (int i = first; i < end; i += step) {
      const float4 v0 = ld4<kSmem>(sv.triangle + 3 * i);
      const float4 e1 = ld4<kSmem>(sv.triangle + 3 * i + 1);
      const float4 e2 = ld4<kSmem>(sv.triangle + 3 * i + 2);
      float t;
      if (triangle_hit_t(r, v3(v0.x, v0.y, v0.z), v3(e1.x, e1.y, e1.z), v3(e2.x, e2.y, e2.z), kTMin, best.t, t))
        consider_le(sc, best, t, make_id(G_TRIANGLE, i));
    }
  } else if (g.type == G_BOX) {
    for (int i = first; i < end; i += step) {
      const float4 p0 = ld4<kSmem>(sv.box + 2 * i);
      const float4 p1 = ld4<kSmem>(sv.box + 2 * i + 1);
      float t, ra, rb;
      if (box_hit_t(r, v3(p0.x, p0.y, p0.z), v3(p1.x, p1.y, p1.z), kTMin, best.t, t, ra, rb) >= 0)
        consider_le(sc, best, t, make_id(G_BOX, i));
    }
  }
}

// render.hpp:30-51 for one ray per TEAM: `member` in [0, team_size) takes every team_size-th object
// of each group.  `act`: this lane's team carries a real ray (the others ride along so that the
// warp stays converged on the shared loads and the shuffles).
template <bool kSmem>
PT_DEV Best closest_hit(const SceneDesc& sc, const SceneView& sv, const Ray& r, Rng& rng, bool act, int member,
                        int team_size) {
  Best best { kInf, -1 };
  const float a = vdot(r.d, r.d);  // sphere.hpp:69, loop invariant
  CullRay cr;
  const int cull_set = make_cull_ray(sc, r, cr);
  const float4* sphere_boxes = sv.sphere_box + cull_set * 2 * (int)sc.n_sphere_chunks;
  const float4* moving_boxes = sv.moving_box + cull_set * 2 * (int)sc.n_moving_chunks;
  const int n_groups = (int)sc.n_groups;
  for (int gi = 0; gi < n_groups; ++gi) {
    const Group g = sv.groups[gi];
    const int end = g.begin + g.count;
    switch (g.type) {
      case G_SPHERE:
        scan_spheres<kSmem, false>(sc, sv.sphere, sphere_boxes, sc.sphere_aux, g.begin, end, team_size, r, a, 0.f, cr, act,
                                   G_SPHERE, best);
        break;
      case G_MOVING_SPHERE:
        scan_spheres<kSmem, true>(sc, sv.moving, moving_boxes, sc.moving_aux, g.begin, end, team_size, r, a,
                                  fdiv(fsub(r.tm, g.time0), g.den), cr, act, G_MOVING_SPHERE, best);
        break;
      case G_MEDIUM: {
        // G_MEDIUM sees the running closest hit of EVERY lower-index object (merge first) and commits
        // unconditionally; every member replays the RNG draw on its replica of the generator.
        if (team_size > 1) team_merge(sc, best, team_size);
        if (act) {
          float t;
          if (medium_hit_t(sc.media[g.begin], r, kTMin, best.t, rng, t)) best.t = t, best.id = make_id(G_MEDIUM, g.begin);
        }
        break;
      }
      default:
        if (act) scan_flat_group<kSmem>(sc, sv, g, r, g.begin + member, team_size, best);
        break;
    }
  }
  if (team_size > 1) team_merge(sc, best, team_size);
  return best;
}

// ---------------------------------------------------------------- shading
struct HitRec {  // hitable.hpp:8-18
  V3 p, normal;
  bool front_face;
  float u, v;
  bool has_uv;
};

// hitable.hpp:20-23
PT_DEV void set_face_normal(HitRec& rec, const Ray& r, V3 outward) {
  rec.front_face = vdot(r.d, outward) < 0.f;
  rec.normal = rec.front_face ? outward : vsub(v3(0.f, 0.f, 0.f), outward);
}

// sphere.hpp:13-24
PT_DEV void mercator(V3 p, float& u, float& v) {
  const float phi = t_atan2(p.z, p.x);
  const float theta = t_asin(p.y);
  u = fsub(1.f, fdiv(fadd(phi, kPi), fmul(2.f, kPi)));
  v = fdiv(fadd(theta, fdiv(kPi, 2.f)), kPi);
}

// texture.hpp:25 / 42-49 / 135-151
PT_DEV V3 texture_value(const SceneDesc& sc, int tex, const HitRec& rec) {
  const pt_texture* t = reinterpret_cast<const pt_texture*>(sc.textures) + tex;
  const int kind = t->kind;
  if (kind == PT_TEX_SOLID) return vld(t->color0);
  if (kind == PT_TEX_CHECKER) {
    const float sines = fmul(fmul(t_sin(fmul(10.f, rec.p.x)), t_sin(fmul(10.f, rec.p.y))), t_sin(fmul(10.f, rec.p.z)));
    return (sines < 0.f) ? vld(t->color0) : vld(t->color1);
  }
  const unsigned long long width = t->width, height = t->height;
  const float fu = fmul(t_fmod1(fmul(rec.u, t->freq)), (float)(width - 1ull));
  const float fv = fmul(fsub(1.f, t_fmod1(fmul(rec.v, t->freq))), (float)(height - 1ull));
  unsigned long long i = (unsigned long long)fu;  // truncation, texture.hpp:139-143
  unsigned long long j = (unsigned long long)fv;
  unsigned long long pix = j * width + i + t->offset;
  if (pix >= sc.n_texture_texels) pix = sc.n_texture_texels - 1ull;  // the reference would read out of bounds
  const unsigned char* td = sc.texture_bytes + pix * 3ull;
  const float scale = fdiv(1.f, 255.f);
  return v3(fmul((float)td[0], scale), fmul((float)td[1], scale), fmul((float)td[2], scale));
}

// material.hpp:62-66
PT_DEV float reflectance(float cosine, float ref_idx) {
  float r0 = fdiv(fsub(1.f, ref_idx), fadd(1.f, ref_idx));
  r0 = fmul(r0, r0);
  return fadd(r0, fmul(fsub(1.f, r0), t_pow5(fsub(1.f, cosine))));
}

// Rebuild the hit_record of the scan winner (the reference fills it inside
// hit(); only the accepted one survives, render.hpp:44-47).
PT_DEV int build_record(const SceneDesc& sc, const SceneView& sv, const Ray& r, const Best& best, HitRec& rec,
                        bool smem) {
  const int type = best.id >> kIdShift;
  const int idx = best.id & (int)kIdMask;
  rec.p = ray_at(r, best.t);
  rec.u = 0.f, rec.v = 0.f, rec.has_uv = false;
  switch (type) {
    case G_SPHERE:
    case G_MOVING_SPHERE: {  // sphere.hpp:78-88
      V3 center;
      const SphereAux* aux;
      if (type == G_SPHERE) {
        const float4 s = smem ? sv.sphere[sphere_slot(idx)] : __ldg(sv.sphere + sphere_slot(idx));
        center = v3(s.x, s.y, s.z);
        aux = sc.sphere_aux + idx;
      } else {
        const float4 s = smem ? sv.moving[moving_slot(idx)] : __ldg(sv.moving + moving_slot(idx));
        const float4 v = smem ? sv.moving[moving_slot(idx) + 2 * kSphereChunk] : __ldg(sv.moving + moving_slot(idx) + 2 * kSphereChunk);
        aux = sc.moving_aux + idx;
        center = moving_center(v3(s.x, s.y, s.z), v3(v.x, v.y, v.z), fdiv(fsub(r.tm, aux->time0), aux->den));
      }
      const V3 outward = vdivs(vsub(rec.p, center), aux->radius);
      set_face_normal(rec, r, outward);
      rec.has_uv = true;  // mercator(rec.normal) evaluated lazily, only for image textures
      return aux->material;
    }
    case G_RECT: {  // rectangle.hpp:42-47
      const float4 q0 = smem ? sv.rect[2 * idx] : __ldg(sv.rect + 2 * idx);
      const float4 q1 = smem ? sv.rect[2 * idx + 1] : __ldg(sv.rect + 2 * idx + 1);
      const int axis = __float_as_int(q1.y);
      const AxisSel s = axis_select(r, axis);
      const float a = fadd(s.oa, fmul(best.t, s.da));
      const float b = fadd(s.ob, fmul(best.t, s.db));
      rec.u = fdiv(fsub(a, q0.x), fsub(q0.y, q0.x));
      rec.v = fdiv(fsub(b, q0.z), fsub(q0.w, q0.z));
      const V3 n = axis == PT_AXIS_XY ? v3(0.f, 0.f, 1.f) : axis == PT_AXIS_XZ ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
      set_face_normal(rec, r, n);
      return sc.rect_aux[idx].material;
    }
    case G_TRIANGLE: {  // triangle.hpp:94-98 (u, v are not written by the reference)
      const TriAux* aux = sc.tri_aux + idx;
      set_face_normal(rec, r, v3(aux->nx, aux->ny, aux->nz));
      return aux->material;
    }
    case G_BOX: {  // box.hpp:29-50: replay the six sides to find the winning one
      const float4 p0 = smem ? sv.box[2 * idx] : __ldg(sv.box + 2 * idx);
      const float4 p1 = smem ? sv.box[2 * idx + 1] : __ldg(sv.box + 2 * idx + 1);
      float t, a, b;
      const V3 lo = v3(p0.x, p0.y, p0.z), hi = v3(p1.x, p1.y, p1.z);
      const int side = box_hit_t(r, lo, hi, kTMin, kInf, t, a, b);
      float a0, a1, b0, b1;
      V3 n;
      if (side < 2) {
        a0 = lo.x, a1 = hi.x, b0 = lo.y, b1 = hi.y, n = v3(0.f, 0.f, 1.f);
      } else if (side < 4) {
        a0 = lo.x, a1 = hi.x, b0 = lo.z, b1 = hi.z, n = v3(0.f, 1.f, 0.f);
      } else {
        a0 = lo.y, a1 = hi.y, b0 = lo.z, b1 = hi.z, n = v3(1.f, 0.f, 0.f);
      }
      rec.u = fdiv(fsub(a, a0), fsub(a1, a0));
      rec.v = fdiv(fsub(b, b0), fsub(b1, b0));
      set_face_normal(rec, r, n);
      return sc.box_aux[idx].material;
    }
    default: {  // constant_medium.hpp:72-76
      rec.normal = v3(1.f, 0.f, 0.f);
      rec.front_face = true;
      return sc.media[idx].material;
    }
  }
}

PT_DEV V3 textured(const SceneDesc& sc, int tex, HitRec& rec) {
  const pt_texture* t = reinterpret_cast<const pt_texture*>(sc.textures) + tex;
  if (t->kind == PT_TEX_IMAGE && rec.has_uv) {
    mercator(rec.normal, rec.u, rec.v);  // sphere.hpp:88
    rec.has_uv = false;
  }
  return texture_value(sc, tex, rec);
}

}  // namespace

// render.hpp:96-99 + camera.hpp:93-100: one camera sample for pixel (px, py); 5 RNG draws
PT_DEV void camera_ray(const pt_camera& cam, int px, int py, float fwidth, float fheight, Rng& rng, Ray& ray) {
  const float u = fdiv(fadd((float)px, rng_float(rng)), fwidth);
  const float v = fdiv(fadd((float)py, rng_float(rng)), fheight);
  float dx, dy;
  rng_in_unit_disk(rng, dx, dy);
  const V3 rd = v3(fmul(cam.lens_radius, dx), fmul(cam.lens_radius, dy), fmul(cam.lens_radius, 0.f));
  const V3 cu = vld(cam.u), cv = vld(cam.v);
  const V3 offset = vadd(v3(fmul(cu.x, rd.x), fmul(cu.y, rd.x), fmul(cu.z, rd.x)),
                         v3(fmul(cv.x, rd.y), fmul(cv.y, rd.y), fmul(cv.z, rd.y)));
  const V3 origin = vld(cam.origin);
  ray.o = vadd(origin, offset);
  ray.d = vsub(vsub(vadd(vadd(vld(cam.lower_left_corner), vscale(u, vld(cam.horizontal))),
                         vscale(v, vld(cam.vertical))),
                    origin),
               offset);
  ray.tm = rng_range(rng, cam.time0, cam.time1);
}

// One iteration of get_color's depth loop after the closest-hit scan (render.hpp:58-91): sky,
// emission or scatter.  Returns true when the path ends; `contribution` is what it adds to the pixel.
PT_DEV bool shade(const SceneDesc& sc, const SceneView& sv, int depth, bool smem, const Best& best, Ray& ray,
                  Rng& rng, V3& att, int& bounce, V3& contribution) {
  contribution = v3(0.f, 0.f, 0.f);
  if (best.id < 0) {
    // background gradient, render.hpp:83-87
    const V3 ud = unit_vector(ray.d);
    const float hit_pt = fmul(0.5f, fadd(ud.y, 1.0f));
    const float w0 = fsub(1.0f, hit_pt);
    const V3 c = vadd(v3(fmul(w0, 1.0f), fmul(w0, 1.0f), fmul(w0, 1.0f)),
                      v3(fmul(hit_pt, 0.5f), fmul(hit_pt, 0.7f), fmul(hit_pt, 1.0f)));
    contribution = vmul(att, c);
    return true;
  }
  HitRec rec;
  const int mat_index = build_record(sc, sv, ray, best, rec, smem);
  const pt_material* m = reinterpret_cast<const pt_material*>(sc.materials) + mat_index;
  const int kind = m->kind;
  bool scattered_ok = true;
  Ray scattered;
  scattered.o = rec.p;
  scattered.tm = ray.tm;
  if (kind == PT_MAT_LAMBERTIAN) {  // material.hpp:18-28
    scattered.d = vadd(rec.normal, rng_unit_vec(rng));
    att = vmul(att, textured(sc, m->texture, rec));
  } else if (kind == PT_MAT_METAL) {  // material.hpp:39-48
    const V3 reflected = reflect(unit_vector(ray.d), rec.normal);
    scattered.d = vadd(reflected, vscale(m->param, rng_in_unit_ball(rng)));
    att = vmul(att, vld(m->albedo));
    scattered_ok = vdot(scattered.d, rec.normal) > 0.f;
  } else if (kind == PT_MAT_DIELECTRIC) {  // material.hpp:68-88
    att = vmul(att, vld(m->albedo));
    const float ref_idx = m->param;
    const float refraction_ratio = rec.front_face ? fdiv(1.0f, ref_idx) : ref_idx;
    const V3 unit_direction = unit_vector(ray.d);
    const float cos_theta = fminf(-vdot(unit_direction, rec.normal), 1.0f);
    const float sin_theta = fsqrt(fsub(1.0f, fmul(cos_theta, cos_theta)));
    const bool cannot_refract = fmul(refraction_ratio, sin_theta) > 1.0f;
    // short-circuit: the RNG is only drawn when refraction is possible
    if (cannot_refract || reflectance(cos_theta, refraction_ratio) > rng_float(rng))
      scattered.d = reflect(unit_direction, rec.normal);
    else
      scattered.d = refract(unit_direction, rec.normal, refraction_ratio);
  } else if (kind == PT_MAT_LIGHTSOURCE) {  // material.hpp:104-108
    contribution = textured(sc, m->texture, rec);  // emitted, NOT attenuated (render.hpp:73)
    scattered_ok = false;
  } else {  // isotropic, material.hpp:119-126
    scattered.d = rng_in_unit_ball(rng);
    att = vmul(att, textured(sc, m->texture, rec));
  }
  if (!scattered_ok) return true;  // render.hpp:73 (emitted is zero for everything but lights)
  ray = scattered;
  ++bounce;
  return bounce == depth;  // render.hpp:91: out of depth -> black
}

// Queue position -> pixel of the region (false: the position falls outside the region and is skipped).
//   order_mode 1  tiles sorted by probed cost, heaviest first (longest-processing-time-first: the
//                 deep pixels of the image start at once, the cheapest ones fill the end of the frame)
//   order_mode 0  consecutive positions spread over the image by a multiplicative permutation
//   order_mode 2  the cost probe itself: every kProbeStep-th pixel of every kProbeStep-th row
PT_DEV bool queue_pixel(const RenderParams& p, unsigned long long pos, int& px, int& py, float*& out_px) {
  unsigned long long k, xx;
  if (p.order_mode == 1) {
    const int tile = p.tile_order[pos / (unsigned long long)(kTile * kTile)];
    const int i = (int)(pos % (unsigned long long)(kTile * kTile));
    xx = (unsigned long long)((tile % p.tiles_x) * kTile + (i % kTile));
    k = (unsigned long long)((tile / p.tiles_x) * kTile + (i / kTile));
    if (xx >= (unsigned long long)p.region.w || k >= (unsigned long long)p.region.h) return false;
  } else if (p.order_mode == 2) {
    const unsigned long long pw = (unsigned long long)((p.region.w + kProbeStep - 1) / kProbeStep);
    xx = (pos % pw) * kProbeStep, k = (pos / pw) * kProbeStep;
  } else {
    const unsigned long long n_pixels = (unsigned long long)p.region.w * (unsigned long long)p.region.h;
    const unsigned long long i = (pos * p.scramble) % n_pixels;
    k = i / (unsigned long long)p.region.w, xx = i - k * (unsigned long long)p.region.w;
  }
  px = p.region.x0 + (int)xx;
  py = p.region.y0 + (int)k * p.region.y_stride;
  out_px = p.out + (long long)k * p.out_row_pitch + 3ll * (long long)xx;
  return true;
}

// ---------------------------------------------------------------- the lane loop
// One warp, k = 32 / team_size pixels at a time, the path state of each pixel replicated in the
// registers of its team (see the header comment); pixels come from the pixel queue.
template <bool kSmem>
PT_DEV void lane_loop(const RenderParams& p, const SceneDesc& sc, const SceneView& sv, int team_size0,
                      unsigned int& n_scans) {
  const pt_camera& cam = p.cam;
  const float fwidth = (float)p.width, fheight = (float)p.height, fspp = (float)p.spp;

  // per-lane path state, replicated across the lanes of a team
  const int lane = (int)(threadIdx.x & 31u);
  int team_size = team_size0;         // lanes per pixel (power of two); grows when the warp is re-packed
  int member = lane & (team_size - 1);
  bool live = false;                  // owns a pixel
  bool need_path = true;              // must start a new camera sample
  bool exhausted_queue = false;       // the pixel queue has run dry
  int px = 0, py = 0;                 // global pixel coordinates
  float* out_px = nullptr;
  int sample = p.spp;
  int bounce = 0;
  Rng rng { 0u };
  Ray ray { v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f), 0.f };
  V3 att = v3(1.f, 1.f, 1.f);
  V3 acc = v3(0.f, 0.f, 0.f);
  int pix_scans = 0;  // closest-hit scans spent on the current pixel

  for (;;) {
    // ---- (R) re-pack: the queue is dry for this warp and at most half of its teams still own a
    // pixel -> move the survivors into teams twice (or more) as large.
    if (team_size < kSphereChunk) {
      const unsigned can_fetch = __ballot_sync(0xffffffffu, !live && !exhausted_queue);
      const unsigned leaders = __ballot_sync(0xffffffffu, live && member == 0);
      const int k_live = __popc(leaders);
      if (can_fetch == 0u && k_live > 0 && 2 * k_live * team_size <= 32) {
        int new_size = team_size;
        while (2 * k_live * new_size <= 32 && new_size < kSphereChunk) new_size <<= 1;
        const int new_team = lane / new_size;
        const bool keep = new_team < k_live;
        const int src = keep ? (int)__fns(leaders, 0u, new_team + 1) : lane;  // leader lane of the new_team-th live team
#define PT_MOVE(x) x = __shfl_sync(0xffffffffu, x, src)
        PT_MOVE(px), PT_MOVE(py), PT_MOVE(sample), PT_MOVE(bounce), PT_MOVE(rng.s), PT_MOVE(pix_scans);
        PT_MOVE(ray.o.x), PT_MOVE(ray.o.y), PT_MOVE(ray.o.z), PT_MOVE(ray.d.x), PT_MOVE(ray.d.y), PT_MOVE(ray.d.z);
        PT_MOVE(ray.tm), PT_MOVE(att.x), PT_MOVE(att.y), PT_MOVE(att.z), PT_MOVE(acc.x), PT_MOVE(acc.y), PT_MOVE(acc.z);
        unsigned long long optr = (unsigned long long)out_px;
        PT_MOVE(optr);
        out_px = (float*)optr;
        int np = need_path ? 1 : 0;
        PT_MOVE(np);
        need_path = np != 0;
#undef PT_MOVE
        live = keep;
        exhausted_queue = true;
        team_size = new_size;
        member = lane & (team_size - 1);
      }
    }

    // ---- (A) path regeneration: render.hpp:94-105 sample loop, :130-133 seeding
    if (need_path && live && sample == p.spp) {
      // final_color /= samples; fb[y][x] = final_color (render.hpp:102-105); one writer per team
      const V3 fin = vdivs(acc, fspp);
      if (member == 0) out_px[0] = fin.x, out_px[1] = fin.y, out_px[2] = fin.z;
      live = false;
    }
    {
      const bool wants = need_path && !live && !exhausted_queue;
      if (__any_sync(0xffffffffu, wants)) {
        unsigned long long idx = 0ull;
        if (wants && member == 0) {  // the team leader pulls the next pixel (skipping tile positions outside the region)
          int tx, ty;
          float* tp;
          do idx = atomicAdd(p.pixel_counter, 1ull);
          while (idx < p.n_positions && !queue_pixel(p, idx, tx, ty, tp));
        }
        idx = __shfl_sync(0xffffffffu, idx, lane - member);
        if (wants) {
          if (idx < p.n_positions) {
            queue_pixel(p, idx, px, py, out_px);
            // std::hash<size_t> is the identity; LocalPseudoRNG takes a uint32_t (rtweekend.hpp:35)
            rng.s = (uint32_t)((unsigned long long)py * (unsigned long long)p.width + (unsigned long long)px);
            acc = v3(0.f, 0.f, 0.f);
            sample = 0;
            pix_scans = 0;
            live = true;
          } else {
            exhausted_queue = true;
            if (p.counters && member == 0) atomicMin(p.counters + 2, globaltimer_ns());  // timeline: queue ran dry
          }
        }
      }
    }
    if (need_path && live) {
      camera_ray(cam, px, py, fwidth, fheight, rng, ray);
      att = v3(1.f, 1.f, 1.f);
      bounce = 0;
      need_path = false;
    }
    if (!__any_sync(0xffffffffu, live)) {
      if (__all_sync(0xffffffffu, exhausted_queue)) break;
      continue;
    }

    // ---- (B) closest hit: render.hpp:60 -> :30-51
    const Best best = closest_hit<kSmem>(sc, sv, ray, rng, live, member, team_size);

    // ---- (C) shade: render.hpp:58-91 (every member of a team computes the same thing)
    if (live) {
      if (member == 0) ++n_scans;
      ++pix_scans;
      V3 contribution;
      if (shade(sc, sv, p.depth, kSmem, best, ray, rng, att, bounce, contribution)) {
        acc = vadd(acc, contribution);
        ++sample;
        need_path = true;
      }
    }
    // A warp that finds itself holding one of the image's deepest pixels stops taking new pixels: as its
    // other pixels finish it is re-packed into ever larger teams, until all 32 lanes scan for the deep
    // pixel and its remaining thousands of bounces take microseconds each instead of a full round.
    if (__any_sync(0xffffffffu, live && pix_scans > kDeepBase + kDeepRate * sample)) exhausted_queue = true;
  }

}

// ---------------------------------------------------------------- the kernel
template <bool kSmem>
__global__ void __launch_bounds__(kBlockThreads, kMinBlocksPerSM) render_kernel(const RenderParams p) {
  extern __shared__ __align__(16) unsigned char smem_blob[];
  __shared__ __align__(8) uint64_t stage_bar;

  const SceneDesc& sc = p.scene;
  const unsigned char* blob_base = sc.blob;
  if constexpr (kSmem) {
    // One TMA bulk copy of the scan blob per CTA; every warp then reads it with
    // broadcast LDS.128 for the rest of the kernel.
    if (threadIdx.x == 0) mbar_init(&stage_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&stage_bar, sc.blob_bytes);
      constexpr uint32_t kPiece = 32768;
      for (uint32_t off = 0; off < sc.blob_bytes; off += kPiece) {
        const uint32_t n = min(kPiece, sc.blob_bytes - off);
        bulk_g2s(smem_blob + off, sc.blob + off, n, &stage_bar);
      }
    }
    mbar_wait(&stage_bar, 0);
    blob_base = smem_blob;
  }
  SceneView sv;
  sv.groups = reinterpret_cast<const Group*>(blob_base + sc.off_groups);
  sv.sphere = reinterpret_cast<const float4*>(blob_base + sc.off_sphere);
  sv.moving = reinterpret_cast<const float4*>(blob_base + sc.off_moving);
  sv.rect = reinterpret_cast<const float4*>(blob_base + sc.off_rect);
  sv.triangle = reinterpret_cast<const float4*>(blob_base + sc.off_triangle);
  sv.box = reinterpret_cast<const float4*>(blob_base + sc.off_box);
  sv.sphere_box = reinterpret_cast<const float4*>(blob_base + sc.off_sphere_box);
  sv.moving_box = reinterpret_cast<const float4*>(blob_base + sc.off_moving_box);

  if (p.counters && threadIdx.x == 0 && blockIdx.x == 0) atomicMin(p.counters + 1, globaltimer_ns());
  unsigned int n_scans = 0;
  lane_loop<kSmem>(p, sc, sv, p.team_size, n_scans);

  if (p.counters && (threadIdx.x & 31) == 0) atomicMax(p.counters + 3, globaltimer_ns());  // timeline: warp retired
  // work counters: one atomic per warp
  unsigned int warp_scans = n_scans;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_scans += __shfl_xor_sync(0xffffffffu, warp_scans, o);
  if ((threadIdx.x & 31) == 0 && p.counters) atomicAdd(p.counters, (unsigned long long)warp_scans);
}

// ---------------------------------------------------------------- the wavefront kernel
// Bulk-synchronous wavefront inside one CTA per SM.  The path state of up to kWavePool pixels lives
// in a structure-of-arrays RAY POOL in shared memory instead of in the registers of fixed lanes, and
// the CTA alternates between phases separated by __syncthreads():
//
//   BOXES   one thread per ray: which sphere chunks does the ray cross (chunk culling, above)?  Every
//           (ray, chunk) pair becomes a work ITEM in a CTA-wide list.
//   SPHERES one thread per ITEM: the 16 spheres of the chunk against the ray; a hit is merged into the
//           ray's winner with one 64-bit shared-memory atomicMin on {t, original object index} -- exactly
//           the reference's rule for spheres (smallest t, then the earlier object, sphere.hpp:77,93).
//           The work per item is the same whatever the ray, so the lanes of a warp stay busy although
//           their rays cross different numbers of chunks, and a CTA with few rays left still has
//           (rays x chunks) items to spread over its threads.
//   LATE    one thread per ray: the groups from the first constant_medium on (flat objects and media in object
//           order against the running closest hit), then what has to happen next -- background, or the kind of
//           the hit material -- and the ray joins that kind's list ("compact divergent material work").
//   SHADE   one warp per unit of up to 32 rays of ONE kind, so the lanes of a warp run the same material
//           code; finished paths start their pixel's next sample (the RNG stream of a pixel is strictly
//           serial) or write the pixel and pull a new one from the global pixel queue.
// A scene with spheres BEHIND a constant_medium in the object list is scanned sequentially per ray in
// BOXES instead (the medium needs the running closest hit of everything before it, and what follows
// needs the medium's).
//
// HAND-OFF QUEUE.  A pixel's samples are serial and a full round takes tens of microseconds, so the
// few pixels that hold ten times the average work (paths bouncing dozens of times inside glass:
// 3 000 scans where the mean is 260) would sit on a critical path longer than the whole frame, and
// at the end of the frame every CTA would drain its own leftovers alone.  A pixel whose scan rate
// marks it as HEAVY is therefore handed, with its complete path state, to a global queue.  It is
// taken over by a CTA that runs SHORT rounds because it keeps only kExpressPool rays in flight: one
// of a few EXPRESS CTAs that do nothing else, or any CTA whose own pixels have run out -- which also
// balances the end of the frame across the whole GPU.  With few rays the phases switch to finer
// work units (a ray's boxes in blocks of 8, a chunk's spheres in quarters).
//
// Results are bit-identical to the lane kernel: the same device functions are called on the same
// per-pixel state, only the assignment of work to lanes differs.
#ifndef PT_WAVE_THREADS
#define PT_WAVE_THREADS 896
#endif
#ifndef PT_WAVE_BLOCKS_PER_SM
#define PT_WAVE_BLOCKS_PER_SM 1
#endif
#ifndef PT_WAVE_ROUNDS
#define PT_WAVE_ROUNDS 1
#endif
#ifndef PT_HEAVY_RATE
#define PT_HEAVY_RATE 10
#endif
#ifndef PT_HEAVY_RATE_DRY
#define PT_HEAVY_RATE_DRY 10
#endif
#ifndef PT_EXPRESS_POOL
#define PT_EXPRESS_POOL 64
#endif
#ifndef PT_WAVE_ITEMS
#define PT_WAVE_ITEMS 4096
#endif
constexpr int kWaveThreads = PT_WAVE_THREADS;
constexpr int kWavePool = PT_WAVE_ROUNDS * kWaveThreads;  // pixels (rays) a CTA keeps in flight: whole scan passes
constexpr int kWaveKinds = 6;                              // 0 = background, 1 + PT_MAT_* otherwise
constexpr int kHeavyRate = PT_HEAVY_RATE;                  // heavy: more than kHeavyBase + rate * samples scans so far
constexpr int kHeavyRateDry = PT_HEAVY_RATE_DRY;           // ... a lower bar once the pixel queue is dry (load sharing)
constexpr int kHeavyBase = 64;
constexpr int kExpressPool = PT_EXPRESS_POOL;              // rays in flight in a CTA that serves the hand-off queue
constexpr int kWaveItems = PT_WAVE_ITEMS;                  // (ray, chunk) items per round; the overflow is scanned in place
constexpr int kWaveItemsStatic = kWaveItems * 3 / 8;       // ... of static spheres (from the front of the list)
constexpr int kWaveItemsMoving = kWaveItems - kWaveItemsStatic;  // ... of moving spheres (from the back)
constexpr unsigned long long kNoHit64 = 0x7f800000ffffffffull;  // {t = +inf, no object}
#ifndef PT_FINE_RAYS
#define PT_FINE_RAYS 160
#endif
constexpr int kFineRays = PT_FINE_RAYS;  // at most this many rays in the round: finer work units (short rounds)
#ifndef PT_FINE_BOXES
#define PT_FINE_BOXES 8
#endif
constexpr int kFineBoxes = PT_FINE_BOXES;            // ... BOXES: a ray's chunks in blocks of this many
constexpr int kFineQuarter = 4;          // ... SPHERES: a chunk's spheres in runs of this many
constexpr int kMaxBoxBlocks = 64;
constexpr int kMaxFlats = 256;
static_assert(kWavePool <= 1024, "an item packs the pool slot into 10 bits");

struct WavePool {
  unsigned long long best64[kWavePool];  // SPHERES: {float bits of t, original object index} of the ray's winner
  uint2 items[kWaveItems];               // {slot | chunk << 10, f bits}: static-sphere items from the front, moving from the back
  float ox[kWavePool], oy[kWavePool], oz[kWavePool], dx[kWavePool], dy[kWavePool], dz[kWavePool], tm[kWavePool];
  float hit_t[kWavePool];
  int hit_id[kWavePool];
  float att_x[kWavePool], att_y[kWavePool], att_z[kWavePool];
  float acc_x[kWavePool], acc_y[kWavePool], acc_z[kWavePool];
  uint32_t rng[kWavePool];
  uint32_t pix[kWavePool];  // position of the pixel in the work queue
  int sample[kWavePool];
  int bounce[kWavePool];
  int scans[kWavePool];               // closest-hit scans spent on the current pixel (< 0: taken over, never handed off again)
  unsigned short list_a[kWavePool];   // rays to scan (unordered)
  unsigned short list_k[kWaveKinds][kWavePool];  // the same rays by kind (what happens next), filled by LATE
  unsigned short free_list[kWavePool];  // hand-off service: pool slots without a pixel
  int counts[8];
  int n_next;      // length of list_a being built
  int n_own;       // of those, pixels this CTA pulled from the pixel queue itself
  int n_items_s, n_items_m;  // static / moving items reserved this round (may exceed what fits)
  int free_count;
  int pixel_dry;   // the pixel queue has run dry
  int service;     // hand-off service: 0 = keep polling, 1 = every producer is done and the queue is empty
  int n_blocks;    // fine BOXES: blocks of <= kFineBoxes chunks over all sphere groups (0: too many, coarse only)
  int4 blocks[kMaxBoxBlocks];  // {group, first chunk, chunks, 0}
  int n_flats;     // rectangles, triangles and boxes in FRONT of the first constant_medium: tested one thread per (ray, object)
  int first_late_group;  // the first constant_medium's group (n_groups if none): from here on the scan is sequential per ray
  int2 flats[kMaxFlats];  // {group type, element}
};

PT_DEV int material_of(const SceneDesc& sc, int id) {
  const int idx = id & (int)kIdMask;
  switch (id >> kIdShift) {
    case G_SPHERE: return sc.sphere_aux[idx].material;
    case G_MOVING_SPHERE: return sc.moving_aux[idx].material;
    case G_RECT: return sc.rect_aux[idx].material;
    case G_TRIANGLE: return sc.tri_aux[idx].material;
    case G_BOX: return sc.box_aux[idx].material;
    default: return sc.media[idx].material;
  }
}

// A ray's winner as one 64-bit word whose UNSIGNED ORDER is the winner rule of pt_packed.h (smaller t, then larger
// key): {float bits of t (t >= 0), 0x7fffffff - key}.  Spheres (key = -1 - object) give codes 0x80000000 + object,
// the others (key = object) 0x7fffffff - object.  A NaN t orders behind +inf and is never taken.
PT_DEV unsigned long long pack_winner(float t, int key) {
  return ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(0x7fffffffu - (uint32_t)key);
}
PT_DEV unsigned long long pack_sphere_winner(const SphereAux* aux, const Best& b) {
  return pack_winner(b.t, aux[b.id & (int)kIdMask].key);
}
PT_DEV Best unpack_winner(const SceneDesc& sc, unsigned long long v) {
  Best best;
  const uint32_t code = (uint32_t)v;
  best.t = __uint_as_float((uint32_t)(v >> 32));
  best.id = code == 0xffffffffu ? -1 : sc.object_id[code >= 0x80000000u ? code - 0x80000000u : 0x7fffffffu - code];
  return best;
}

template <bool kSmem>
__global__ void __launch_bounds__(kWaveThreads, PT_WAVE_BLOCKS_PER_SM) render_wave_kernel(const RenderParams p) {
  extern __shared__ __align__(16) unsigned char smem_blob[];
  __shared__ __align__(8) uint64_t stage_bar;
  __shared__ SceneDesc staged_scene;  // the scene descriptor with the staged tables' pointers redirected to shared memory

  // dynamic shared memory: the ray pool first (at a compile-time offset, so that its accesses need no address arithmetic),
  // then the staged part of the arena
  constexpr uint32_t kPoolBytes = ((uint32_t)sizeof(WavePool) + 127u) & ~127u;
  unsigned char* const smem_stage = smem_blob + kPoolBytes;
  const unsigned char* blob_base = p.scene.blob;
  const uint32_t staged = kSmem ? p.staged_bytes : 0u;  // the scan blob, and the side tables behind it when they fit too
  if constexpr (kSmem) {
    if (threadIdx.x == 0) mbar_init(&stage_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&stage_bar, staged);
      constexpr uint32_t kPiece = 32768;
      for (uint32_t off = 0; off < staged; off += kPiece)
        bulk_g2s(smem_stage + off, p.scene.blob + off, min(kPiece, staged - off), &stage_bar);
    }
    blob_base = smem_stage;
  }
  if (threadIdx.x == 0) {
    staged_scene = p.scene;
    const unsigned char* g0 = p.scene.blob;
    auto redirect = [&](auto& ptr) {
      const size_t off = (size_t)(reinterpret_cast<const unsigned char*>(ptr) - g0);
      if (kSmem && off < (size_t)staged) ptr = reinterpret_cast<decltype(ptr + 0)>(smem_stage + off);
    };
    redirect(staged_scene.sphere_aux), redirect(staged_scene.moving_aux), redirect(staged_scene.rect_aux);
    redirect(staged_scene.tri_aux), redirect(staged_scene.box_aux), redirect(staged_scene.media);
    redirect(staged_scene.keys), redirect(staged_scene.object_id);
    const unsigned char* mats = static_cast<const unsigned char*>(staged_scene.materials);
    redirect(mats);
    staged_scene.materials = mats;
  }
  __syncthreads();
  if constexpr (kSmem) mbar_wait(&stage_bar, 0);
  const SceneDesc& sc = staged_scene;
  WavePool& W = *reinterpret_cast<WavePool*>(smem_blob);
  SceneView sv;
  sv.groups = reinterpret_cast<const Group*>(blob_base + sc.off_groups);
  sv.sphere = reinterpret_cast<const float4*>(blob_base + sc.off_sphere);
  sv.moving = reinterpret_cast<const float4*>(blob_base + sc.off_moving);
  sv.rect = reinterpret_cast<const float4*>(blob_base + sc.off_rect);
  sv.triangle = reinterpret_cast<const float4*>(blob_base + sc.off_triangle);
  sv.box = reinterpret_cast<const float4*>(blob_base + sc.off_box);
  sv.sphere_box = reinterpret_cast<const float4*>(blob_base + sc.off_sphere_box);
  sv.moving_box = reinterpret_cast<const float4*>(blob_base + sc.off_moving_box);

  if (p.counters && threadIdx.x == 0 && blockIdx.x == 0) atomicMin(p.counters + 1, globaltimer_ns());
  const unsigned long long t_give_up = globaltimer_ns() + 30000000000ull;  // watchdog against a hung queue
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rot = lane & (kSphereChunk - 1);
  const pt_camera& cam = p.cam;
  const float fwidth = (float)p.width, fheight = (float)p.height, fspp = (float)p.spp;
  const unsigned lane_lt = (1u << lane) - 1u;
  const HeavyQueue& hq = p.heavy;
  const bool express = (int)blockIdx.x < p.n_express;  // this CTA only serves the hand-off queue
  const bool sequential_scan = sc.n_late_sphere_groups != 0u;
  const int n_groups = (int)sc.n_groups;
  unsigned int n_scans = 0;

  // Pull the next pixel of the queue; false (and the CTA-wide flag set) when the queue is dry.
  // The first p.express_positions positions of the queue (the most expensive tiles of the LPT order) belong to the
  // express CTAs, which trace them in short rounds from the start; everybody else begins behind them.
  auto next_pixel = [&](uint32_t& pixq, Rng& rng, int& px, int& py) -> bool {
    for (;;) {
      if (*reinterpret_cast<volatile int*>(&W.pixel_dry)) return false;  // (a set-once flag; a stale 0 only costs one more atomic)
      unsigned long long pos;
      if (express) {
        pos = atomicAdd(p.pixel_counter + 1, 1ull);
        if (pos >= p.express_positions) {
          atomicExch(&W.pixel_dry, 1);
          return false;
        }
      } else {
        pos = p.express_positions + atomicAdd(p.pixel_counter, 1ull);
        if (pos >= p.n_positions) {
          atomicExch(&W.pixel_dry, 1);
          if (p.counters) atomicMin(p.counters + 2, globaltimer_ns());  // timeline: queue ran dry
          return false;
        }
      }
      float* unused;
      if (!queue_pixel(p, pos, px, py, unused)) continue;  // a tile position outside the region
      pixq = (uint32_t)pos;
      // std::hash<size_t> is the identity; LocalPseudoRNG takes a uint32_t (rtweekend.hpp:35)
      rng.s = (uint32_t)((unsigned long long)py * (unsigned long long)p.width + (unsigned long long)px);
      if (p.order_mode == 2) rng.s = (rng.s * 2654435761u) | 1u;  // cost probe: a throw-away stream, never the pixel's
      return true;
    }
  };
  // Take the next entry of the hand-off queue: its index, or -1 when there is none right now.
  auto take_heavy = [&]() -> int {
    unsigned int h = ld_volatile_u32(hq.ctrl + 0);
    for (int attempt = 0; attempt < 8; ++attempt) {
      const unsigned int t = min(ld_volatile_u32(hq.ctrl + 1), hq.cap);
      if (h >= t) return -1;
      const unsigned int seen = atomicCAS(hq.ctrl + 0, h, h + 1u);
      if (seen == h) {
        while (ld_volatile_u32(hq.ready + h) != hq.stamp) {
          if (globaltimer_ns() > t_give_up) {
            if (p.counters) atomicExch(p.counters + 4, 1ull);  // reported as an error by the host
            return -1;
          }
          __nanosleep(100);
        }
        __threadfence();
        return (int)h;
      }
      h = seen;
    }
    return -1;
  };
  // Append the live slots of this warp to the next scan list (one shared-memory atomic per warp).
  auto append = [&](bool alive, bool own, int slot) {
    const unsigned m = __ballot_sync(0xffffffffu, alive);
    const unsigned mo = __ballot_sync(0xffffffffu, alive && own);
    if (m != 0u) {
      int base = 0;
      if (lane == 0) {
        base = atomicAdd(&W.n_next, __popc(m));
        if (mo != 0u) atomicAdd(&W.n_own, __popc(mo));
      }
      base = __shfl_sync(0xffffffffu, base, 0);
      if (alive) W.list_a[base + __popc(m & lane_lt)] = (unsigned short)slot;
    }
  };
  auto store_ray = [&](int slot, const Ray& ray, V3 att, V3 acc, Rng rng, int bounce, int sample) {
    W.ox[slot] = ray.o.x, W.oy[slot] = ray.o.y, W.oz[slot] = ray.o.z;
    W.dx[slot] = ray.d.x, W.dy[slot] = ray.d.y, W.dz[slot] = ray.d.z, W.tm[slot] = ray.tm;
    W.att_x[slot] = att.x, W.att_y[slot] = att.y, W.att_z[slot] = att.z;
    W.acc_x[slot] = acc.x, W.acc_y[slot] = acc.y, W.acc_z[slot] = acc.z;
    W.rng[slot] = rng.s, W.bounce[slot] = bounce, W.sample[slot] = sample;
    W.best64[slot] = kNoHit64;
  };
  auto load_ray = [&](int slot) -> Ray {
    Ray ray;
    ray.o = v3(W.ox[slot], W.oy[slot], W.oz[slot]);
    ray.d = v3(W.dx[slot], W.dy[slot], W.dz[slot]);
    ray.tm = W.tm[slot];
    return ray;
  };
  // BOXES for one ray and the chunks [cb, cb + nb) of one sphere group: every crossed chunk becomes an item;
  // what does not fit into the item list is scanned here and now (`inl`).
  auto emit_items = [&](int slot, const Ray& ray, const CullRay& cr, const float4* boxes, bool moving, int cb, int nb, float f,
                        float a, Best& inl) {
    uint32_t hits = chunk_hits<kSmem>(boxes, cb, nb, cr, kInf);
    if (hits == 0u) return;
    const int cnt = __popc(hits);
    int at = atomicAdd(moving ? &W.n_items_m : &W.n_items_s, cnt);
    // static items grow from the front, moving ones from the back, each within its fixed share of the list
    while (hits) {
      const int top = 31 - __clz((int)hits);
      hits &= ~(1u << top);
      const int chunk = cb + (nb - 1 - top);
      if (at < (moving ? kWaveItemsMoving : kWaveItemsStatic)) {
        W.items[moving ? kWaveItems - 1 - at : at] = make_uint2((uint32_t)slot | ((uint32_t)chunk << 10), __float_as_uint(f));
      } else {
        if (p.counters) atomicAdd(p.counters + 15, 1ull);  // stats: items scanned in place (tests check that it happens)
        if (moving)
          scan_chunk<kSmem, true, kSphereChunk, 1>(sc, sv.moving, sc.moving_aux, chunk, rot, ray, a, filter_a(a), f, G_MOVING_SPHERE, inl);
        else
          scan_chunk<kSmem, false, kSphereChunk, 1>(sc, sv.sphere, sc.sphere_aux, chunk, rot, ray, a, filter_a(a), 0.f, G_SPHERE, inl);
      }
      ++at;
    }
  };

  int mode = 0;  // 0: this CTA's share of the pixel queue (none for an express CTA); 1: hand-off service
  if (tid == 0) {
    int nb = 0;
    for (int gi = 0; gi < n_groups && nb >= 0; ++gi) {
      const Group g = sv.groups[gi];
      if (g.type != G_SPHERE && g.type != G_MOVING_SPHERE) continue;
      const int c_end = (g.begin + g.count) / kSphereChunk;
      for (int cb = g.begin / kSphereChunk; cb < c_end; cb += kFineBoxes) {
        if (nb == kMaxBoxBlocks) {
          nb = -1;
          break;
        }
        W.blocks[nb++] = make_int4(gi, cb, min(kFineBoxes, c_end - cb), 0);
      }
    }
    W.n_blocks = nb < 0 ? 0 : nb;
    // the flat objects in front of the first medium (as many as fit; the rest stays with the sequential part)
    int nf = 0, late = n_groups;
    for (int gi = 0; gi < n_groups; ++gi) {
      const Group g = sv.groups[gi];
      if (g.type == G_MEDIUM || ((g.type == G_RECT || g.type == G_TRIANGLE || g.type == G_BOX) && nf + 6 * g.count > kMaxFlats)) {
        late = gi;
        break;
      }
      if (g.type == G_RECT || g.type == G_TRIANGLE)
        for (int i = 0; i < g.count; ++i) W.flats[nf++] = make_int2(g.type, g.begin + i);
      if (g.type == G_BOX)  // a box is six independent sides (box.hpp:20-25): the closest side is the box's hit
        for (int i = 0; i < g.count; ++i)
          for (int side = 0; side < 6; ++side) W.flats[nf++] = make_int2(G_BOX | (side << 8), g.begin + i);
    }
    W.n_flats = nf, W.first_late_group = late;
  }
  if (tid < 8) W.counts[tid] = 0;
  if (tid == 0) W.n_next = 0, W.n_own = 0, W.n_items_s = 0, W.n_items_m = 0, W.free_count = 0, W.pixel_dry = 0, W.service = 0;
  __syncthreads();
  if (!express) {  // (an express CTA goes straight to the hand-off service, whose first source is its reserved tiles)
    // ---- start: every pool slot (up to this CTA's fair share of the image) takes a pixel
    const int cap = p.pool_cap;
    for (int s0 = warp * 32; s0 < kWavePool; s0 += kWaveThreads) {
      const int slot = s0 + lane;
      bool alive = false;
      if (slot < cap) {
        uint32_t pixq;
        Rng rng;
        int px, py;
        if (next_pixel(pixq, rng, px, py)) {
          Ray ray;
          camera_ray(cam, px, py, fwidth, fheight, rng, ray);
          store_ray(slot, ray, v3(1.f, 1.f, 1.f), v3(0.f, 0.f, 0.f), rng, 0, 0);
          W.pix[slot] = pixq, W.scans[slot] = express ? -1 : 0;  // (an express CTA never hands a pixel off)
          alive = true;
        }
      }
      append(alive, !express, slot);
    }
    __syncthreads();
  }

  for (unsigned int round = 0;; ++round) {
    if (mode == 1) {
      // ---- hand-off service: fill the free pool slots from the global queue (polled every 8th round: a poll is
      // two round trips to L2, and these rounds are the frame's critical path)
      const int room = W.free_count, waiting = W.n_next;
      if (room > 0 && ((round & 7u) == 0u || waiting == 0)) {
        __syncthreads();  // everybody has read the two (and decided alike) before anybody changes them
        int got = -1;
        if (tid < room) {
          // an express CTA's first source is the head of the LPT order (reserved for it), then the hand-off queue -- which
          // it serves from the start, whenever it has room
          uint32_t pixq;
          Rng rng;
          int px, py;
          if (express && next_pixel(pixq, rng, px, py)) {
            const int slot = (int)W.free_list[atomicSub(&W.free_count, 1) - 1];
            Ray ray;
            camera_ray(cam, px, py, fwidth, fheight, rng, ray);
            store_ray(slot, ray, v3(1.f, 1.f, 1.f), v3(0.f, 0.f, 0.f), rng, 0, 0);
            W.pix[slot] = pixq, W.scans[slot] = -1;  // (never handed off again)
            W.list_a[atomicAdd(&W.n_next, 1)] = (unsigned short)slot;
          } else {
            got = take_heavy();
          }
        }
        if (got >= 0) {
          const int slot = (int)W.free_list[atomicSub(&W.free_count, 1) - 1];
          const float* e = hq.entries + (size_t)got * kHeavyEntryWords;
          Ray ray;
          ray.o = v3(__ldcg(e + 4), __ldcg(e + 5), __ldcg(e + 6));
          ray.d = v3(__ldcg(e + 7), __ldcg(e + 8), __ldcg(e + 9));
          ray.tm = __ldcg(e + 10);
          store_ray(slot, ray, v3(__ldcg(e + 11), __ldcg(e + 12), __ldcg(e + 13)), v3(__ldcg(e + 14), __ldcg(e + 15), __ldcg(e + 16)),
                    Rng { __float_as_uint(__ldcg(e + 1)) }, __float_as_int(__ldcg(e + 3)), __float_as_int(__ldcg(e + 2)));
          W.pix[slot] = __float_as_uint(__ldcg(e + 0)), W.scans[slot] = -1;
          W.list_a[atomicAdd(&W.n_next, 1)] = (unsigned short)slot;
          if (p.counters) {  // stats: how long did the pixel wait in the queue
            const unsigned long long waited = (uint32_t)((uint32_t)globaltimer_ns() - __float_as_uint(__ldcg(e + 17)));
            atomicAdd(p.counters + 12, waited), atomicMax(p.counters + 13, waited);
          }
        }
        __syncthreads();
      }
    }
    const int n = W.n_next;  // rays to trace this round
    if (n == 0) {
      if (mode == 0) {
        // ---- no regular work (left): say so, then serve the hand-off queue until the whole GPU is finished
        if (p.order_mode == 2) break;  // the cost probe hands nothing off
        mode = 1;
        for (int s = tid; s < kExpressPool; s += kWaveThreads) W.free_list[s] = (unsigned short)s;
        if (tid == 0) {
          W.free_count = kExpressPool;
          __threadfence();
          atomicAdd(hq.ctrl + 2, 1u);
          if (p.counters && !express) atomicMin(p.counters + 5, globaltimer_ns()), atomicMax(p.counters + 6, globaltimer_ns());
        }
        __syncthreads();
        continue;
      }
      // idle: every producer is done and the queue is empty (or the watchdog fired) => nothing will ever arrive again
      if (tid == 0)
        W.service = ((ld_volatile_u32(hq.ctrl + 2) >= gridDim.x &&
                      ld_volatile_u32(hq.ctrl + 0) >= min(ld_volatile_u32(hq.ctrl + 1), hq.cap)) ||
                     globaltimer_ns() > t_give_up)
                        ? 1
                        : 0;
      __syncthreads();
      if (W.service == 1) break;
      __nanosleep(300);
      continue;  // (the next write of W.service is behind the barrier at the loop top)
    }
    const bool fine = n <= kFineRays && W.n_blocks > 0;
#ifdef PT_PHASE_TIMING
    long long pt_t0 = clock64();
#define PT_PHASE(k)                                                                                   \
  if (tid == 0 && p.counters) {                                                                       \
    const long long now = clock64();                                                                  \
    atomicAdd(p.counters + 16 + (k), (unsigned long long)(now - pt_t0));                              \
    pt_t0 = now;                                                                                      \
  }
#else
#define PT_PHASE(k)
#endif
    if (mode == 1 && tid == 0 && p.counters) atomicAdd(p.counters + 8, 1ull), atomicAdd(p.counters + 9, (unsigned long long)n);  // stats

    // ---- BOXES (or, with media in the scene, the whole sequential scan)
    if (sequential_scan) {
      for (int e = tid; e < n; e += kWaveThreads) {
        const int slot = (int)W.list_a[e];
        const Ray ray = load_ray(slot);
        Rng rng { W.rng[slot] };
        const Best best = closest_hit<kSmem>(sc, sv, ray, rng, true, 0, 1);
        W.hit_t[slot] = best.t, W.hit_id[slot] = best.id;
        W.rng[slot] = rng.s;  // a constant_medium may have drawn from it (constant_medium.hpp:65)
      }
    } else {
      // one work unit = one ray x (all its chunks | a block of kFineBoxes chunks)
      const int n_units = fine ? n * W.n_blocks : n;
      for (int w = tid; w < n_units; w += kWaveThreads) {
        const int e = fine ? w / W.n_blocks : w;
        const int slot = (int)W.list_a[e];
        const Ray ray = load_ray(slot);
        CullRay cr;
        const int cull_set = make_cull_ray(sc, ray, cr);
        const float a = vdot(ray.d, ray.d);  // sphere.hpp:69
        Best inl { kInf, -1 };
        unsigned long long v = kNoHit64;
        const int4 blk = fine ? W.blocks[w - e * W.n_blocks] : make_int4(0, 0, 0, 0);
        for (int gi = fine ? blk.x : 0; gi < (fine ? blk.x + 1 : n_groups); ++gi) {
          const Group g = sv.groups[gi];
          if (g.type != G_SPHERE && g.type != G_MOVING_SPHERE) continue;
          const bool moving = g.type == G_MOVING_SPHERE;
          const float4* boxes = moving ? sv.moving_box + cull_set * 2 * (int)sc.n_moving_chunks
                                       : sv.sphere_box + cull_set * 2 * (int)sc.n_sphere_chunks;
          const float f = moving ? fdiv(fsub(ray.tm, g.time0), g.den) : 0.f;
          // the group's outsized spheres (a ground sphere ...) are tested right here, one by one: their chunks are never
          // culled and mostly padding (the same sphere for every lane: broadcast loads)
          // (short rounds leave them to SPHERES as items of their never-culled chunks: a shorter chain here)
          const int open_chunks = fine ? 0 : (g.n_open + kSphereChunk - 1) / kSphereChunk;
          if (g.n_open > 0 && !fine) {
            const float af = filter_a(a);
            for (int i = g.begin; i < g.begin + g.n_open; ++i) {
              float cx, cy, cz, r2f;
              if (moving) {
                const float4* ps = sv.moving + moving_slot(i);
                if ((int)sphere_filter_bits<kSmem, true>(ps, f, ray, af) < 0) {
                  sphere_center<kSmem, true>(ps, f, cx, cy, cz, r2f);
                  sphere_roots_scan(sc, inl, ray, a, cx, cy, cz, exact_r2(sc.moving_aux, i), make_id(G_MOVING_SPHERE, i));
                }
              } else {
                const float4* ps = sv.sphere + sphere_slot(i);
                if ((int)sphere_filter_bits<kSmem, false>(ps, f, ray, af) < 0) {
                  sphere_center<kSmem, false>(ps, f, cx, cy, cz, r2f);
                  sphere_roots_scan(sc, inl, ray, a, cx, cy, cz, exact_r2(sc.sphere_aux, i), make_id(G_SPHERE, i));
                }
              }
            }
          }
          const int c_group = g.begin / kSphereChunk + open_chunks;  // the first chunk with a box
          const int c_first = fine ? max(blk.y, c_group) : c_group;
          const int c_end = fine ? blk.y + blk.z : (g.begin + g.count) / kSphereChunk;
          if (inl.id >= 0) {
            const unsigned long long w64 = pack_sphere_winner(moving ? sc.moving_aux : sc.sphere_aux, inl);
            if (w64 < v) v = w64;
            inl.t = kInf, inl.id = -1;
          }
          for (int cb = c_first; cb < c_end; cb += 32) {
            emit_items(slot, ray, cr, boxes, moving, cb, min(32, c_end - cb), f, a, inl);
            if (inl.id >= 0) {  // scanned in place: fold into the ray's winner
              const unsigned long long w64 = pack_sphere_winner(moving ? sc.moving_aux : sc.sphere_aux, inl);
              if (w64 < v) v = w64;
              inl.t = kInf, inl.id = -1;
            }
          }
        }
        if (!fine && W.n_flats != 0) {
          // many rays: the flat objects in front of the first medium one after the other, here (few rays: one thread per
          // (ray, object) in SPHERES)
          Best fb { kInf, -1 };
          for (int gi = 0; gi < W.first_late_group; ++gi) {
            const Group g = sv.groups[gi];
            if (g.type == G_RECT || g.type == G_TRIANGLE || g.type == G_BOX) scan_flat_group<kSmem>(sc, sv, g, ray, g.begin, 1, fb);
          }
          if (fb.id >= 0) {
            const unsigned long long w64 = pack_winner(fb.t, key_of(sc, fb.id));
            if (w64 < v) v = w64;
          }
        }
        if (v != kNoHit64) atomicMin(&W.best64[slot], v);
      }
    }
    __syncthreads();
    PT_PHASE(0)

    // ---- SPHERES: one thread per (ray, chunk) item, or per quarter of one
    if (!sequential_scan) {
      const int n_s = min(W.n_items_s, kWaveItemsStatic), n_m = min(W.n_items_m, kWaveItemsMoving);
#ifdef PT_PHASE_TIMING
      if (tid == 0 && p.counters) atomicAdd(p.counters + 23, (unsigned long long)(n_s + n_m));
#endif
      if (!fine) {
        // one index space for both kinds (a thread's items follow each other without a pass boundary in between); the
        // moving items start at a warp boundary so that a warp runs one kind's code
        const int m_base = (n_s + 31) & ~31;
        for (int i = tid; i < m_base + n_m; i += kWaveThreads) {
          if (i >= n_s && i < m_base) continue;
          const bool moving = i >= m_base;
          const uint2 it = W.items[moving ? kWaveItems - 1 - (i - m_base) : i];
          const int slot = (int)(it.x & 1023u);
          const Ray ray = load_ray(slot);
          const float a = vdot(ray.d, ray.d);
          Best b { kInf, -1 };
          if (moving)
            scan_chunk<kSmem, true, kSphereChunk, 1>(sc, sv.moving, sc.moving_aux, (int)(it.x >> 10), rot, ray, a, filter_a(a),
                                                     __uint_as_float(it.y), G_MOVING_SPHERE, b);
          else
            scan_chunk<kSmem, false, kSphereChunk, 1>(sc, sv.sphere, sc.sphere_aux, (int)(it.x >> 10), rot, ray, a, filter_a(a), 0.f,
                                                      G_SPHERE, b);
          if (b.id >= 0) atomicMin(&W.best64[slot], pack_sphere_winner(moving ? sc.moving_aux : sc.sphere_aux, b));
        }
      } else {
        // Short rounds, ONE unit per thread where possible: a quarter of an item (its kParts threads are neighbouring lanes:
        // lane & 15 = 4 m + q takes slots 4 m + q + 4 k, like a team of 4), or one (ray, flat object) pair for the
        // rectangles, triangles and box sides in front of the first medium (object-major: a warp tests one object).
        constexpr int kParts = kSphereChunk / kFineQuarter;
        static_assert(kWaveThreads % kParts == 0 && kParts == 4, "an item's threads must be lanes 4 m .. 4 m + 3");
        const int n_quarters = kParts * (n_s + n_m);
        const int f_base = (n_quarters + 31) & ~31;
        const int n_flats = W.n_flats;
        for (int w = tid; w < f_base + n * n_flats; w += kWaveThreads) {
          if (w < n_quarters) {
            const int i = w / kParts;
            const bool moving = i >= n_s;
            const uint2 it = W.items[moving ? kWaveItems - 1 - (i - n_s) : i];
            const int slot = (int)(it.x & 1023u);
            const Ray ray = load_ray(slot);
            const float a = vdot(ray.d, ray.d);
            Best b { kInf, -1 };
            if (moving)
              scan_chunk<kSmem, true, kFineQuarter, kParts>(sc, sv.moving, sc.moving_aux, (int)(it.x >> 10), rot, ray, a, filter_a(a),
                                                            __uint_as_float(it.y), G_MOVING_SPHERE, b);
            else
              scan_chunk<kSmem, false, kFineQuarter, kParts>(sc, sv.sphere, sc.sphere_aux, (int)(it.x >> 10), rot, ray, a, filter_a(a),
                                                             0.f, G_SPHERE, b);
            if (b.id >= 0) atomicMin(&W.best64[slot], pack_sphere_winner(moving ? sc.moving_aux : sc.sphere_aux, b));
          } else if (w >= f_base) {
            const int j = (w - f_base) / n;
            const int2 fo = W.flats[j];
            const int slot = (int)W.list_a[(w - f_base) - j * n];
            const Ray ray = load_ray(slot);
            Best b { kInf, -1 };
            if ((fo.x & 255) == G_BOX) {
              const float4 p0 = ld4<kSmem>(sv.box + 2 * fo.y);
              const float4 p1 = ld4<kSmem>(sv.box + 2 * fo.y + 1);
              float t, ra, rb;
              if (box_side_hit_t(ray, v3(p0.x, p0.y, p0.z), v3(p1.x, p1.y, p1.z), fo.x >> 8, kTMin, kInf, t, ra, rb))
                b.t = t, b.id = make_id(G_BOX, fo.y);
            } else {
              Group g {};
              g.type = fo.x, g.begin = fo.y, g.count = 1;
              scan_flat_group<kSmem>(sc, sv, g, ray, fo.y, 1, b);
            }
            if (b.id >= 0) atomicMin(&W.best64[slot], pack_winner(b.t, key_of(sc, b.id)));
          }
        }
      }
      __syncthreads();
    }
    PT_PHASE(1)

    // ---- LATE: the groups from the first constant_medium on; what happens next to the ray
    if (tid == 0) W.n_next = 0, W.n_own = 0, W.n_items_s = 0, W.n_items_m = 0;  // (SHADE builds the next round's list)
    for (int e = tid; e < n; e += kWaveThreads) {
      const int slot = (int)W.list_a[e];
      Best best;
      if (sequential_scan) {
        best.t = W.hit_t[slot], best.id = W.hit_id[slot];
      } else {
        best = unpack_winner(sc, W.best64[slot]);
        const int late = W.first_late_group;
        if (late < n_groups) {
          // from the first constant_medium on, in group order against the running closest hit; a medium commits
          // unconditionally and may draw from the pixel's stream (constant_medium.hpp:52-65)
          const Ray ray = load_ray(slot);
          Rng rng { W.rng[slot] };
          for (int gi = late; gi < n_groups; ++gi) {
            const Group g = sv.groups[gi];
            if (g.type == G_RECT || g.type == G_TRIANGLE || g.type == G_BOX) {
              scan_flat_group<kSmem>(sc, sv, g, ray, g.begin, 1, best);
            } else if (g.type == G_MEDIUM) {
              float t;
              if (medium_hit_t(sc.media[g.begin], ray, kTMin, best.t, rng, t)) best.t = t, best.id = make_id(G_MEDIUM, g.begin);
            }
          }
          W.rng[slot] = rng.s;
        }
        W.hit_t[slot] = best.t, W.hit_id[slot] = best.id;
      }
      W.scans[slot] += W.scans[slot] >= 0 ? 1 : -1;  // (a taken-over pixel counts downwards: -1 - rounds in the service)
      int kind = 0;
      if (best.id >= 0) kind = 1 + reinterpret_cast<const pt_material*>(sc.materials)[material_of(sc, best.id)].kind;
      W.list_k[kind][atomicAdd(&W.counts[kind], 1)] = (unsigned short)slot;  // "sorted" by kind as a side effect
      ++n_scans;
    }
    __syncthreads();
    PT_PHASE(2)
    PT_PHASE(3)

    // ---- SHADE: one warp per unit of up to 32 rays of ONE kind (no divergence on the material)
    int unit_base[kWaveKinds + 1], kind_count[kWaveKinds];  // 32-ray shading units in front of each kind
    {
      int units = 0;
#pragma unroll
      for (int k = 0; k < kWaveKinds; ++k) {
        kind_count[k] = W.counts[k];
        unit_base[k] = units, units += (kind_count[k] + 31) >> 5;
      }
      unit_base[kWaveKinds] = units;
    }
    const int heavy_rate = *reinterpret_cast<volatile int*>(&W.pixel_dry) ? kHeavyRateDry : kHeavyRate;
    for (int u = warp; u < unit_base[kWaveKinds]; u += kWaveThreads / 32) {
      int kind = 0, e = 0, e_end = 0;
#pragma unroll
      for (int k = 0; k < kWaveKinds; ++k)
        if (u >= unit_base[k] && u < unit_base[k + 1]) kind = k, e = ((u - unit_base[k]) << 5) + lane, e_end = kind_count[k];
      const bool act = e < e_end;
      const int slot = act ? (int)W.list_k[kind][e] : 0;
      bool alive = false, own = false;
      if (act) {
        Ray ray = load_ray(slot);
        const Best best { W.hit_t[slot], W.hit_id[slot] };
        V3 att = v3(W.att_x[slot], W.att_y[slot], W.att_z[slot]);
        V3 acc = v3(W.acc_x[slot], W.acc_y[slot], W.acc_z[slot]);
        Rng rng { W.rng[slot] };
        int bounce = W.bounce[slot], sample = W.sample[slot];
        uint32_t pixq = W.pix[slot];
        const int scans = W.scans[slot];
        own = scans >= 0;
        V3 contribution;
        bool new_pixel = false;
        alive = true;
        if (shade(sc, sv, p.depth, kSmem, best, ray, rng, att, bounce, contribution)) {
          // the path ended: render.hpp:100-105
          acc = vadd(acc, contribution);
          int px, py;
          float* out_px;
          queue_pixel(p, pixq, px, py, out_px);
          if (++sample == p.spp) {
            if (p.order_mode == 2) {
              p.probe_cost[pixq] = scans;  // cost probe: how deep did one sample of this pixel go
            } else {
              const V3 fin = vdivs(acc, fspp);
              out_px[0] = fin.x, out_px[1] = fin.y, out_px[2] = fin.z;
            }
            new_pixel = true;
          } else {
            camera_ray(cam, px, py, fwidth, fheight, rng, ray);
            att = v3(1.f, 1.f, 1.f);
            bounce = 0;
          }
        }
        // a heavy pixel leaves for a CTA that runs short rounds, with its complete state
        if (!new_pixel && own && p.order_mode != 2 && scans > kHeavyBase + heavy_rate * sample &&
            ld_volatile_u32(hq.ctrl + 1) < hq.cap) {
          const unsigned int i = atomicAdd(hq.ctrl + 1, 1u);
          if (i < hq.cap) {
            float* q = hq.entries + (size_t)i * kHeavyEntryWords;
            __stcg(q + 0, __uint_as_float(pixq)), __stcg(q + 1, __uint_as_float(rng.s));
            __stcg(q + 2, __int_as_float(sample)), __stcg(q + 3, __int_as_float(bounce));
            __stcg(q + 4, ray.o.x), __stcg(q + 5, ray.o.y), __stcg(q + 6, ray.o.z);
            __stcg(q + 7, ray.d.x), __stcg(q + 8, ray.d.y), __stcg(q + 9, ray.d.z), __stcg(q + 10, ray.tm);
            __stcg(q + 11, att.x), __stcg(q + 12, att.y), __stcg(q + 13, att.z);
            __stcg(q + 14, acc.x), __stcg(q + 15, acc.y), __stcg(q + 16, acc.z);
            __stcg(q + 17, __uint_as_float((uint32_t)globaltimer_ns()));
            __threadfence();
            *reinterpret_cast<volatile unsigned int*>(hq.ready + i) = hq.stamp;
            new_pixel = true;
          }
        }
        if (new_pixel) {
          if (!own && p.counters) atomicMax(p.counters + 14, (unsigned long long)(-1 - scans));  // stats: longest stay in the service
          int px, py;
          alive = mode == 0 && next_pixel(pixq, rng, px, py);
          if (alive) {
            camera_ray(cam, px, py, fwidth, fheight, rng, ray);
            att = v3(1.f, 1.f, 1.f);
            acc = v3(0.f, 0.f, 0.f);
            bounce = 0, sample = 0;
            W.pix[slot] = pixq, W.scans[slot] = express ? -1 : 0;
            own = !express;
          } else if (mode == 1) {
            W.free_list[atomicAdd(&W.free_count, 1)] = (unsigned short)slot;  // refilled from the hand-off queue
          }
        }
        if (alive) store_ray(slot, ray, att, acc, rng, bounce, sample);
      }
      append(alive, own, slot);
    }
    __syncthreads();
    if (tid < 8) W.counts[tid] = 0;  // (everybody has read them; LATE of the next round is two barriers away)
    PT_PHASE(4)
#ifdef PT_PHASE_TIMING
    if (tid == 0 && p.counters) atomicAdd(p.counters + 21, 1ull), atomicAdd(p.counters + 22, (unsigned long long)n);
#endif
  }

  if (p.counters && lane == 0) atomicMax(p.counters + 3, globaltimer_ns());  // timeline: warp retired
  unsigned int warp_scans = n_scans;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_scans += __shfl_xor_sync(0xffffffffu, warp_scans, o);
  if (lane == 0 && p.counters) atomicAdd(p.counters, (unsigned long long)warp_scans);
}

// ---------------------------------------------------------------- LPT tile order
// One block: bin the tiles by probed cost (sum of the probes inside the tile), then list them from the
// most expensive bin to the cheapest (counting sort; the order inside a bin does not matter).
constexpr int kCostBins = 1024;
__global__ void __launch_bounds__(1024) tile_order_kernel(const int* __restrict__ probe_cost, int region_w, int region_h,
                                                           int tiles_x, int tiles_y, int* __restrict__ tile_order,
                                                           int* __restrict__ tile_bin) {
  __shared__ int hist[kCostBins];
  __shared__ int start[kCostBins];
  const int n_tiles = tiles_x * tiles_y;
  const int pw = (region_w + kProbeStep - 1) / kProbeStep, ph = (region_h + kProbeStep - 1) / kProbeStep;
  for (int b = threadIdx.x; b < kCostBins; b += blockDim.x) hist[b] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    const int tx = t % tiles_x, ty = t / tiles_x;
    int cost = 0;
    for (int j = 0; j < kTile / kProbeStep; ++j)
      for (int i = 0; i < kTile / kProbeStep; ++i) {
        const int qx = tx * (kTile / kProbeStep) + i, qy = ty * (kTile / kProbeStep) + j;
        if (qx < pw && qy < ph) cost += probe_cost[qy * pw + qx];
      }
    const int bin = min(cost, kCostBins - 1);
    tile_bin[t] = bin;
    atomicAdd(&hist[bin], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = kCostBins - 1; b >= 0; --b) start[b] = run, run += hist[b];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) tile_order[atomicAdd(&start[tile_bin[t]], 1)] = t;
}

cudaError_t launch_tile_order(const int* probe_cost, int region_w, int region_h, int tiles_x, int tiles_y,
                              int* tile_order, int* scratch, cudaStream_t stream) {
  tile_order_kernel<<<1, 1024, 0, stream>>>(probe_cost, region_w, region_h, tiles_x, tiles_y, tile_order, scratch);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- launch
int max_smem_blob_bytes(int device) {
  int optin = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  return optin - 1024;
}

cudaError_t launch_render(const RenderParams& p, int device, int grid_override, cudaStream_t stream,
                          LaunchInfo* info) {
  int sms = 0;
  cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (err != cudaSuccess) return err;
  RenderParams q = p;
  const unsigned long long pixels = (unsigned long long)p.region.w * (unsigned long long)p.region.h;
  if (p.kernel_kind == 0) {
    // ---- wavefront kernel: one CTA per SM, ray pool + (when it fits) the scan blob in shared memory
    // shared memory: the ray pool, and in front of it the scan blob with the side tables (else the blob alone, else nothing)
    const size_t pool_bytes = sizeof(WavePool);
    auto with_pool = [&](size_t bytes) { return (pool_bytes + 127u) / 128u * 128u + bytes; };
    const long long most = (long long)max_smem_blob_bytes(device) - (long long)sizeof(SceneDesc);
    q.staged_bytes = (long long)with_pool(p.scene.stage_bytes) <= most  ? p.scene.stage_bytes
                     : (long long)with_pool(p.scene.blob_bytes) <= most ? p.scene.blob_bytes
                                                                        : 0u;
    const bool smem = q.staged_bytes != 0u;
    const size_t dyn = smem ? with_pool(q.staged_bytes) : pool_bytes;
    auto kernel = smem ? render_wave_kernel<true> : render_wave_kernel<false>;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (err != cudaSuccess) return err;
    // every CTA must be resident at once: the express warps wait for all CTAs to report
    int resident = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, kWaveThreads, dyn);
    if (err != cudaSuccess) return err;
    if (resident < 1) return cudaErrorLaunchOutOfResources;
    if (resident > PT_WAVE_BLOCKS_PER_SM) resident = PT_WAVE_BLOCKS_PER_SM;
    int grid = grid_override > 0 ? grid_override : sms * resident;
    if (grid > sms * resident) grid = sms * resident;
    const unsigned long long share = (pixels + (unsigned long long)grid - 1ull) / (unsigned long long)grid;
    q.pool_cap = (int)(share < 32ull ? 32ull : (share > (unsigned long long)kWavePool ? (unsigned long long)kWavePool : share));
    // a few CTAs only serve the hand-off queue (short rounds for the deepest pixels of the image)
    q.n_express = p.n_express >= 0 ? p.n_express : (grid >= 64 ? (grid * 23 + 100) / 200 : 0);  // 17 of 148: measured best on the default scene
    if (q.n_express >= grid) q.n_express = grid - 1;
    if (p.order_mode == 2) {
      q.n_express = 0;
      q.n_positions = (unsigned long long)((p.region.w + kProbeStep - 1) / kProbeStep) *
                      (unsigned long long)((p.region.h + kProbeStep - 1) / kProbeStep);
    } else if (p.order_mode == 1) {
      q.n_positions = (unsigned long long)p.tiles_x * (unsigned long long)p.tiles_y * (unsigned long long)(kTile * kTile);
    } else {
      q.n_positions = pixels;
    }
    // with the LPT order the express CTAs start on the most expensive tiles, one pixel per pool slot
    q.express_positions = 0;
    if (p.order_mode == 1) {
      q.express_positions = (unsigned long long)q.n_express * (unsigned long long)kExpressPool;
      if (q.express_positions > q.n_positions) q.express_positions = q.n_positions;
    }
    {
      const unsigned long long sh = (q.n_positions + (unsigned long long)grid - 1ull) / (unsigned long long)grid;
      q.pool_cap = (int)(sh < 32ull ? 32ull : (sh > (unsigned long long)kWavePool ? (unsigned long long)kWavePool : sh));
    }
    // pixel-order permutation pos -> (pos * scramble) mod pixels: a multiplier near pixels / golden ratio,
    // made coprime with the pixel count so that it is a bijection
    unsigned long long mul = (unsigned long long)((double)pixels * 0.6180339887498949) | 1ull;
    auto gcd = [](unsigned long long a, unsigned long long b) {
      while (b) {
        const unsigned long long t = a % b;
        a = b, b = t;
      }
      return a;
    };
    while (pixels > 1 && gcd(mul % pixels, pixels) != 1ull) mul += 2ull;
    q.scramble = pixels > 1 ? mul % pixels : 1ull;
    if (q.scramble == 0ull) q.scramble = 1ull;
    if (info) info->grid = grid, info->block = kWaveThreads, info->smem_bytes = (int)dyn, info->blocks_per_sm = 1, info->staged = smem, info->team_size = 0;
    kernel<<<grid, kWaveThreads, dyn, stream>>>(q);
    return cudaGetLastError();
  }
  // ---- lane kernel: a pixel per lane team, state in registers
  if (p.order_mode == 1)
    q.n_positions = (unsigned long long)p.tiles_x * (unsigned long long)p.tiles_y * (unsigned long long)(kTile * kTile);
  else
    q.n_positions = pixels, q.order_mode = 0, q.scramble = 1ull;  // row-major
  const bool smem = (int)p.scene.blob_bytes <= max_smem_blob_bytes(device);
  const size_t dyn = smem ? p.scene.blob_bytes : 0;
  auto kernel = smem ? render_kernel<true> : render_kernel<false>;
  if (smem) {
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (err != cudaSuccess) return err;
  }
  int per_sm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlockThreads, dyn);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > kMaxBlocksPerSM) per_sm = kMaxBlocksPerSM;
  int grid = grid_override > 0 ? grid_override : sms * per_sm;
  // Lanes per pixel at launch: one in the normal case; with fewer than kMinPixelsPerTeam pixels per
  // team (a small region, or an image strongly scaled over many GPUs) the teams start larger.
  if (q.team_size <= 0) {
    const unsigned long long lanes = (unsigned long long)grid * kBlockThreads;
    int t = 1;
    while (t < kSphereChunk && pixels * (unsigned long long)t * 2ull < lanes * (unsigned long long)kMinPixelsPerTeam) t <<= 1;
    q.team_size = t;
  }
  if (info) info->grid = grid, info->block = kBlockThreads, info->smem_bytes = (int)dyn, info->blocks_per_sm = per_sm, info->staged = smem, info->team_size = q.team_size;
  kernel<<<grid, kBlockThreads, dyn, stream>>>(q);
  return cudaGetLastError();
}

}  // namespace ptb
