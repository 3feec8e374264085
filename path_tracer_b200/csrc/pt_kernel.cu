// pt_kernel.cu -- the per-pixel path-tracing loop of triSYCL/path_tracer
// (reference include/render.hpp:25-106 and everything it calls) as ONE
// persistent sm_100a kernel.
//
// Execution model (B200-first, not a translation of the SYCL kernel):
//   * Persistent CTAs; every LANE owns one pixel at a time and walks that
//     pixel's `spp` samples serially -- the xorshift32 stream of a pixel is
//     consumed in a data-dependent way (render.hpp:95-101), so samples of one
//     pixel cannot be split.  When a pixel is finished the lane pulls the next
//     pixel index from a global atomic queue ("path regeneration"), so all 32
//     lanes re-converge at the closest-hit scan with live rays until the
//     queue runs dry.
//   * The closest-hit scan (render.hpp:30-51) is the hot loop.  The scene's
//     scan blob (pt_packed.h) is staged into shared memory with one
//     cp.async.bulk (TMA) per CTA and streamed with warp-broadcast LDS.128.
//     Sphere tests are split in two: a branch-free discriminant pass over a
//     chunk of 32 spheres that only records a per-lane candidate bitmask, and
//     an exact root pass (sqrt, IEEE division, range and tie rules) over the
//     few set bits.  Only `t` and the object id are tracked; the full
//     hit_record (point, normal, face, u, v) is rebuilt once, for the winner.
//   * Shading (material scatter / emission / textures) is divergent by
//     nature; it is short compared with the scan and runs per lane.
// All arithmetic follows the operation order of the reference; see
// pt_device.cuh for the numerics contract.
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
#define PT_SCAN_UNROLL 4
#endif
constexpr int kScanUnroll = PT_SCAN_UNROLL;  // spheres per hot-loop trip
#ifndef PT_BOOST_ROUNDS
#define PT_BOOST_ROUNDS 0
#endif
#ifndef PT_HEAVY_PER_SAMPLE
#define PT_HEAVY_PER_SAMPLE 8
#endif
constexpr int kBoostRounds = PT_BOOST_ROUNDS;        // short rounds for heavy pixels between two normal rounds
constexpr int kHeavyPerSample = PT_HEAVY_PER_SAMPLE;  // a pixel is heavy above this many scans per sample ...
constexpr int kHeavyBase = 32;                        // ... plus this head start
#ifndef PT_RAMP_BETA
#define PT_RAMP_BETA 0.0f
#endif
constexpr float kRampBeta = PT_RAMP_BETA;  // ramp-down starts when remaining pixels < 32 * warps * beta
#ifndef PT_TEAM_MAX_LIVE
#define PT_TEAM_MAX_LIVE 16
#endif
constexpr int kTeamMaxLive = PT_TEAM_MAX_LIVE;  // at or below this many live lanes the warp scans in teams

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

PT_DEV int key_of(const SceneDesc& sc, int id) {
  return sc.keys[sc.key_base[id >> kIdShift] + (uint32_t)(id & (int)kIdMask)];
}

// Winner rule: minimum t, then maximum key (pt_packed.h).  Called with a
// candidate that already satisfies its own primitive's range test.
PT_DEV void consider(const SceneDesc& sc, Best& best, float t, int id) {
  if (t < best.t) {
    best.t = t, best.id = id;
  } else if (t == best.t) {
    if (best.id < 0 || key_of(sc, id) > key_of(sc, best.id)) best.t = t, best.id = id;
  }
}
// rect / triangle / box accept with `!(t > max)`, which lets NaN through
// (rectangle.hpp:36, triangle.hpp:91): mirror that.
PT_DEV void consider_le(const SceneDesc& sc, Best& best, float t, int id) {
  if (t == best.t) {
    if (best.id < 0 || key_of(sc, id) > key_of(sc, best.id)) best.id = id;
  } else {
    best.t = t, best.id = id;
  }
}

// ---------------------------------------------------------------- primitives
// Exact roots of one sphere for the scan (sphere.hpp:74-105 with max = +inf;
// the running-closest filter is applied by consider()).  `a` = dot(d,d).
PT_DEV void sphere_roots_scan(const SceneDesc& sc, Best& best, const Ray& r, float a, float cx, float cy,
                              float cz, float r2, int id) {
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
PT_DEV int box_hit_t(const Ray& r, V3 p0, V3 p1, float tmin, float tmax, float& t_out, float& a_out,
                     float& b_out) {
  int side = -1;
  float closest = tmax;
#pragma unroll 1
  for (int s = 0; s < 6; ++s) {
    const int axis = s >> 1;  // PT_AXIS_XY, PT_AXIS_XZ, PT_AXIS_YZ
    const bool hi = (s & 1) == 0;
    float a0, a1, b0, b1, k;
    if (axis == PT_AXIS_XY)
      a0 = p0.x, a1 = p1.x, b0 = p0.y, b1 = p1.y, k = hi ? p1.z : p0.z;
    else if (axis == PT_AXIS_XZ)
      a0 = p0.x, a1 = p1.x, b0 = p0.z, b1 = p1.z, k = hi ? p1.y : p0.y;
    else
      a0 = p0.y, a1 = p1.y, b0 = p0.z, b1 = p1.z, k = hi ? p1.x : p0.x;
    float t, a, b;
    if (rect_hit_t(r, axis, a0, a1, b0, b1, k, tmin, closest, t, a, b))
      side = s, closest = t, t_out = t, a_out = a, b_out = b;
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
// render.hpp:30-51.  `live` lanes carry a real ray; the others ride along so
// that the warp stays converged on the broadcast loads.
template <bool kSmem>
PT_DEV Best closest_hit(const SceneDesc& sc, const SceneView& sv, const Ray& r, Rng& rng, bool live) {
  Best best { kInf, -1 };
  const float a = vdot(r.d, r.d);  // sphere.hpp:69, loop invariant
  const int n_groups = (int)sc.n_groups;
  for (int gi = 0; gi < n_groups; ++gi) {
    const Group g = sv.groups[gi];
    switch (g.type) {
      case G_SPHERE: {
#pragma unroll 1
        for (int base = g.begin; base < g.begin + g.count; base += kSphereChunk) {
          const float4* __restrict__ p = sv.sphere + base;
          uint32_t mask = 0;
          // Branch-free discriminant pass; kScanUnroll spheres per loop trip keeps
          // the hot loop inside the instruction cache (DESIGN.md "scan loop").
#pragma unroll 1
          for (int it = 0; it < kSphereChunk; it += kScanUnroll) {
            uint32_t nib = 0;
#pragma unroll
            for (int j = 0; j < kScanUnroll; ++j) {
              const float4 s = ld4<kSmem>(p + it + j);
              const float ocx = fsub(r.o.x, s.x), ocy = fsub(r.o.y, s.y), ocz = fsub(r.o.z, s.z);
              const float b = fadd(fadd(fmul(ocx, r.d.x), fmul(ocy, r.d.y)), fmul(ocz, r.d.z));
              const float c = fsub(fadd(fadd(fmul(ocx, ocx), fmul(ocy, ocy)), fmul(ocz, ocz)), s.w);
              const float disc = fsub(fmul(b, b), fmul(a, c));
              if (disc > 0.f) nib |= (1u << j);
            }
            mask |= nib << it;
          }
          if (!live) mask = 0;
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const float4 s = ld4<kSmem>(p + j);
            sphere_roots_scan(sc, best, r, a, s.x, s.y, s.z, s.w, make_id(G_SPHERE, base + j));
          }
        }
        break;
      }
      case G_MOVING_SPHERE: {
        const float f = fdiv(fsub(r.tm, g.time0), g.den);  // sphere.hpp:55
#pragma unroll 1
        for (int base = g.begin; base < g.begin + g.count; base += kSphereChunk) {
          const float4* __restrict__ p = sv.moving + 2 * base;
          uint32_t mask = 0;
#pragma unroll 1
          for (int it = 0; it < kSphereChunk; it += kScanUnroll) {
            uint32_t nib = 0;
#pragma unroll
            for (int j = 0; j < kScanUnroll; ++j) {
              const float4 s = ld4<kSmem>(p + 2 * (it + j));
              const float4 v = ld4<kSmem>(p + 2 * (it + j) + 1);
              const float cx = fadd(s.x, fmul(f, v.x)), cy = fadd(s.y, fmul(f, v.y)), cz = fadd(s.z, fmul(f, v.z));
              const float ocx = fsub(r.o.x, cx), ocy = fsub(r.o.y, cy), ocz = fsub(r.o.z, cz);
              const float b = fadd(fadd(fmul(ocx, r.d.x), fmul(ocy, r.d.y)), fmul(ocz, r.d.z));
              const float c = fsub(fadd(fadd(fmul(ocx, ocx), fmul(ocy, ocy)), fmul(ocz, ocz)), s.w);
              const float disc = fsub(fmul(b, b), fmul(a, c));
              if (disc > 0.f) nib |= (1u << j);
            }
            mask |= nib << it;
          }
          if (!live) mask = 0;
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const float4 s = ld4<kSmem>(p + 2 * j);
            const float4 v = ld4<kSmem>(p + 2 * j + 1);
            const float cx = fadd(s.x, fmul(f, v.x)), cy = fadd(s.y, fmul(f, v.y)), cz = fadd(s.z, fmul(f, v.z));
            sphere_roots_scan(sc, best, r, a, cx, cy, cz, s.w, make_id(G_MOVING_SPHERE, base + j));
          }
        }
        break;
      }
      case G_RECT: {
        if (live)
          for (int i = g.begin; i < g.begin + g.count; ++i) {
            const float4 q0 = ld4<kSmem>(sv.rect + 2 * i);
            const float4 q1 = ld4<kSmem>(sv.rect + 2 * i + 1);
            float t, ra, rb;
            if (rect_hit_t(r, __float_as_int(q1.y), q0.x, q0.y, q0.z, q0.w, q1.x, kTMin, best.t, t, ra, rb))
              consider_le(sc, best, t, make_id(G_RECT, i));
          }
        break;
      }
      case G_TRIANGLE: {
        if (live)
          for (int i = g.begin; i < g.begin + g.count; ++i) {
            const float4 v0 = ld4<kSmem>(sv.triangle + 3 * i);
            const float4 e1 = ld4<kSmem>(sv.triangle + 3 * i + 1);
            const float4 e2 = ld4<kSmem>(sv.triangle + 3 * i + 2);
            float t;
            if (triangle_hit_t(r, v3(v0.x, v0.y, v0.z), v3(e1.x, e1.y, e1.z), v3(e2.x, e2.y, e2.z), kTMin, best.t, t))
              consider_le(sc, best, t, make_id(G_TRIANGLE, i));
          }
        break;
      }
      case G_BOX: {
        if (live)
          for (int i = g.begin; i < g.begin + g.count; ++i) {
            const float4 p0 = ld4<kSmem>(sv.box + 2 * i);
            const float4 p1 = ld4<kSmem>(sv.box + 2 * i + 1);
            float t, ra, rb;
            if (box_hit_t(r, v3(p0.x, p0.y, p0.z), v3(p1.x, p1.y, p1.z), kTMin, best.t, t, ra, rb) >= 0)
              consider_le(sc, best, t, make_id(G_BOX, i));
          }
        break;
      }
      default: {  // G_MEDIUM: sees the running closest of every lower-index object, commits unconditionally
        if (live) {
          float t;
          if (medium_hit_t(sc.media[g.begin], r, kTMin, best.t, rng, t)) best.t = t, best.id = make_id(G_MEDIUM, g.begin);
        }
        break;
      }
    }
  }
  return best;
}


// ---------------------------------------------------------------- team scan (tail mode)
// When only k < 32 lanes of a warp still own a pixel (the work queue has run
// dry), one ray per lane would leave most of the warp idle and -- worse --
// leave the deepest paths of the image on a serial critical path.  The warp
// is then split into teams of T = 32 / pow2ceil(k) lanes, one team per live
// ray: member m of a team tests objects m, m+T, ... exactly, and the members'
// winners are merged with the same (minimum t, maximum key) rule, so the
// result is bit-identical to the per-lane scan for any T.  A constant_medium
// still sees the running closest hit of ALL lower-index objects: the team
// merges before evaluating it, and every member replays its RNG draw on a copy
// of the owner's generator.
struct KeyedBest {
  float t;
  int id;
  int key;
};

PT_DEV void offer(KeyedBest& m, float t, int id, int key) {
  if (t < m.t || (t == m.t && key > m.key) || m.id < 0) m.t = t, m.id = id, m.key = key;
}
// rect / triangle / box let NaN through their range test (rectangle.hpp:36)
PT_DEV void offer_le(KeyedBest& m, float t, int id, int key) {
  if (!(t == m.t) || key > m.key || m.id < 0) m.t = t, m.id = id, m.key = key;
}

PT_DEV void team_merge(KeyedBest& m, int team_size) {
  for (int o = team_size >> 1; o > 0; o >>= 1) {
    const float ot = __shfl_xor_sync(0xffffffffu, m.t, o);
    const int oid = __shfl_xor_sync(0xffffffffu, m.id, o);
    const int okey = __shfl_xor_sync(0xffffffffu, m.key, o);
    if (oid >= 0 && (m.id < 0 || ot < m.t || (ot == m.t && okey > m.key))) m.t = ot, m.id = oid, m.key = okey;
  }
}

PT_DEV void team_sphere(KeyedBest& m, const Ray& r, float a, float cx, float cy, float cz, float r2, int id,
                        const SphereAux* aux) {
  const float ocx = fsub(r.o.x, cx), ocy = fsub(r.o.y, cy), ocz = fsub(r.o.z, cz);
  const float b = fadd(fadd(fmul(ocx, r.d.x), fmul(ocy, r.d.y)), fmul(ocz, r.d.z));
  const float c = fsub(fadd(fadd(fmul(ocx, ocx), fmul(ocy, ocy)), fmul(ocz, ocz)), r2);
  const float disc = fsub(fmul(b, b), fmul(a, c));
  if (!(disc > 0.f) || (b > 0.f && c > 0.f)) return;
  const float sq = fsqrt(disc);
  const float t0 = fdiv(fsub(-b, sq), a);
  if (t0 < kInf && t0 > kTMin) {
    if (t0 <= m.t || m.id < 0) offer(m, t0, id, aux->key);
    return;
  }
  const float t1 = fdiv(fadd(-b, sq), a);
  if (t1 < kInf && t1 > kTMin && (t1 <= m.t || m.id < 0)) offer(m, t1, id, aux->key);
}

// `member` in [0, team_size); `active` = this team carries a live ray.
template <bool kSmem>
PT_DEV Best team_closest_hit(const SceneDesc& sc, const SceneView& sv, const Ray& r, Rng& rng, int member,
                             int team_size, bool active) {
  KeyedBest m { kInf, -1, 0 };
  const float a = vdot(r.d, r.d);
  const int n_groups = (int)sc.n_groups;
  for (int gi = 0; gi < n_groups; ++gi) {
    const Group g = sv.groups[gi];
    const int end = active ? g.begin + g.count : 0;
    switch (g.type) {
      case G_SPHERE: {
#pragma unroll 1
        for (int i = g.begin + member; i < end; i += team_size) {
          const float4 s = ld4<kSmem>(sv.sphere + i);
          team_sphere(m, r, a, s.x, s.y, s.z, s.w, make_id(G_SPHERE, i), sc.sphere_aux + i);
        }
        break;
      }
      case G_MOVING_SPHERE: {
        const float f = fdiv(fsub(r.tm, g.time0), g.den);
#pragma unroll 1
        for (int i = g.begin + member; i < end; i += team_size) {
          const float4 s = ld4<kSmem>(sv.moving + 2 * i);
          const float4 v = ld4<kSmem>(sv.moving + 2 * i + 1);
          const float cx = fadd(s.x, fmul(f, v.x)), cy = fadd(s.y, fmul(f, v.y)), cz = fadd(s.z, fmul(f, v.z));
          team_sphere(m, r, a, cx, cy, cz, s.w, make_id(G_MOVING_SPHERE, i), sc.moving_aux + i);
        }
        break;
      }
      case G_RECT: {
#pragma unroll 1
        for (int i = g.begin + member; i < end; i += team_size) {
          const float4 q0 = ld4<kSmem>(sv.rect + 2 * i);
          const float4 q1 = ld4<kSmem>(sv.rect + 2 * i + 1);
          float t, ra, rb;
          if (rect_hit_t(r, __float_as_int(q1.y), q0.x, q0.y, q0.z, q0.w, q1.x, kTMin, m.t, t, ra, rb))
            offer_le(m, t, make_id(G_RECT, i), sc.rect_aux[i].key);
        }
        break;
      }
      case G_TRIANGLE: {
#pragma unroll 1
        for (int i = g.begin + member; i < end; i += team_size) {
          const float4 v0 = ld4<kSmem>(sv.triangle + 3 * i);
          const float4 e1 = ld4<kSmem>(sv.triangle + 3 * i + 1);
          const float4 e2 = ld4<kSmem>(sv.triangle + 3 * i + 2);
          float t;
          if (triangle_hit_t(r, v3(v0.x, v0.y, v0.z), v3(e1.x, e1.y, e1.z), v3(e2.x, e2.y, e2.z), kTMin, m.t, t))
            offer_le(m, t, make_id(G_TRIANGLE, i), sc.tri_aux[i].key);
        }
        break;
      }
      case G_BOX: {
#pragma unroll 1
        for (int i = g.begin + member; i < end; i += team_size) {
          const float4 p0 = ld4<kSmem>(sv.box + 2 * i);
          const float4 p1 = ld4<kSmem>(sv.box + 2 * i + 1);
          float t, ra, rb;
          if (box_hit_t(r, v3(p0.x, p0.y, p0.z), v3(p1.x, p1.y, p1.z), kTMin, m.t, t, ra, rb) >= 0)
            offer_le(m, t, make_id(G_BOX, i), sc.box_aux[i].key);
        }
        break;
      }
      default: {
        team_merge(m, team_size);  // every member now holds the running closest hit of all lower-index objects
        float t;
        if (active && medium_hit_t(sc.media[g.begin], r, kTMin, m.id < 0 ? kInf : m.t, rng, t))
          m.t = t, m.id = make_id(G_MEDIUM, g.begin), m.key = sc.media[g.begin].key;
        break;
      }
    }
  }
  team_merge(m, team_size);
  return Best { m.id < 0 ? kInf : m.t, m.id };
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
        const float4 s = smem ? sv.sphere[idx] : __ldg(sv.sphere + idx);
        center = v3(s.x, s.y, s.z);
        aux = sc.sphere_aux + idx;
      } else {
        const float4 s = smem ? sv.moving[2 * idx] : __ldg(sv.moving + 2 * idx);
        const float4 v = smem ? sv.moving[2 * idx + 1] : __ldg(sv.moving + 2 * idx + 1);
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

  if (p.counters && threadIdx.x == 0 && blockIdx.x == 0) atomicMin(p.counters + 1, globaltimer_ns());
  const pt_camera& cam = p.cam;
  const unsigned long long n_pixels = (unsigned long long)p.region.w * (unsigned long long)p.region.h;
  const float fwidth = (float)p.width, fheight = (float)p.height, fspp = (float)p.spp;

  // per-lane path state
  bool live = false;          // owns a pixel
  bool need_path = true;      // must start a new camera sample
  int px = 0, py = 0;         // global pixel coordinates
  float* out_px = nullptr;
  int sample = p.spp;         // == spp forces a pixel fetch first
  int bounce = 0;
  Rng rng { 0u };
  Ray ray { v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f), 0.f };
  V3 att = v3(1.f, 1.f, 1.f);
  V3 acc = v3(0.f, 0.f, 0.f);
  unsigned int n_scans = 0;
  unsigned int pix_scans = 0;  // closest-hit scans spent on the current pixel
  bool exhausted_queue = false;
  int warp_cap = 32;  // pixels this warp may hold (ramp-down at the end of the queue)
  const unsigned long long ramp_div =
      max(1ull, (unsigned long long)((float)(gridDim.x * (blockDim.x >> 5)) * kRampBeta));

  for (unsigned iter = 0;; ++iter) {
    // Boost rounds: the image's deepest pixels (paths bouncing dozens of times inside glass) hold
    // ten times the average work and, one bounce per full-cost round, would sit on a serial
    // critical path longer than the whole frame.  Between two normal rounds the warp therefore
    // runs kBoostRounds short rounds for its HEAVY pixels only, scanned by lane teams.
    const bool boost = (iter % (unsigned)(kBoostRounds + 1)) != 0u;
    const bool heavy = live && pix_scans > (unsigned)(kHeavyBase + kHeavyPerSample * sample);
    if (boost && !__any_sync(0xffffffffu, heavy)) continue;
    const bool part = !boost || heavy;  // this lane takes part in this round

    // ---- (A) path regeneration: render.hpp:94-105 sample loop, :130-133 seeding
    // A pixel is finished when its last sample ended: write it out (render.hpp:102-105).
    if (need_path && live && part && sample == p.spp) {
      const V3 fin = vdivs(acc, fspp);
      out_px[0] = fin.x, out_px[1] = fin.y, out_px[2] = fin.z;
      live = false;
    }
    // Pixel fetch with ramp-down: while plenty of pixels remain every lane holds one (cap 32).
    // Towards the end of the queue a warp may only hold `cap` pixels, cap halving as the queue
    // drains, so that the pixels still in flight are scanned by ever larger lane teams (shorter
    // per-pixel latency) instead of leaving a long tail of nearly idle warps.
    {
      const bool wants = need_path && !live && !exhausted_queue && !boost;
      const unsigned want_mask = __ballot_sync(0xffffffffu, wants);
      if (want_mask != 0u) {
        const unsigned busy_mask = __ballot_sync(0xffffffffu, live);
        const int room = warp_cap - __popc(busy_mask);
        const int my_rank = __popc(want_mask & ((1u << (threadIdx.x & 31u)) - 1u));
        unsigned fetched_pos = 0u;
        if (wants && my_rank < room) {
          const unsigned long long idx = atomicAdd(p.pixel_counter, 1ull);
          if (idx < n_pixels) {
            const int k = (int)(idx / (unsigned long long)p.region.w);
            const int xx = (int)(idx - (unsigned long long)k * (unsigned long long)p.region.w);
            px = p.region.x0 + xx;
            py = p.region.y0 + k * p.region.y_stride;
            out_px = p.out + (long long)k * p.out_row_pitch + 3ll * xx;
            // std::hash<size_t> is the identity; LocalPseudoRNG takes a uint32_t (rtweekend.hpp:35)
            rng.s = (uint32_t)((unsigned long long)py * (unsigned long long)p.width + (unsigned long long)px);
            acc = v3(0.f, 0.f, 0.f);
            sample = 0;
            pix_scans = 0u;
            live = true;
            fetched_pos = (unsigned)min(idx + 1ull, 0xffffffffull);
          } else {
            exhausted_queue = true;
            fetched_pos = 0xffffffffu;
            if (p.counters) atomicMin(p.counters + 2, globaltimer_ns());  // timeline: queue ran dry
          }
        }
        const unsigned seen = __reduce_max_sync(0xffffffffu, fetched_pos);
        if (seen != 0u) {
          const unsigned long long remaining = seen >= n_pixels ? 0ull : n_pixels - seen;
          // pixels a warp may hold = remaining pixels per warp (scaled), rounded down to a power of two
          const unsigned long long per_warp = remaining / ramp_div;
          warp_cap = per_warp >= 32ull ? 32 : (per_warp <= 1ull ? 1 : (1 << (31 - __clz((int)per_warp))));
        }
      }
    }
    if (need_path && live && part) {
        // render.hpp:96-99 + camera.hpp:93-100
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
        att = v3(1.f, 1.f, 1.f);
        bounce = 0;
        need_path = false;
    }
    if (!boost && !__any_sync(0xffffffffu, live)) break;
    const bool act = live && part;
    const unsigned live_mask = __ballot_sync(0xffffffffu, act);  // lanes with a ray to trace this round
    if (live_mask == 0u) continue;

    // ---- (B) closest hit: render.hpp:60 -> :30-51
    Best best { kInf, -1 };
    const int n_live = __popc(live_mask);
    if (n_live <= kTeamMaxLive) {
      // few rays: one TEAM of lanes per ray (see team_closest_hit)
      const int lane = (int)(threadIdx.x & 31u);
      int team_size = 32;
      while (team_size * n_live > 32) team_size >>= 1;  // largest power of two with team_size * n_live <= 32
      const int team = lane / team_size;
      const bool active = team < n_live;
      const int owner = active ? (int)__fns(live_mask, 0u, team + 1) : 0;  // lane of the team-th live ray
      Ray rr;
      rr.o.x = __shfl_sync(0xffffffffu, ray.o.x, owner), rr.o.y = __shfl_sync(0xffffffffu, ray.o.y, owner);
      rr.o.z = __shfl_sync(0xffffffffu, ray.o.z, owner), rr.d.x = __shfl_sync(0xffffffffu, ray.d.x, owner);
      rr.d.y = __shfl_sync(0xffffffffu, ray.d.y, owner), rr.d.z = __shfl_sync(0xffffffffu, ray.d.z, owner);
      rr.tm = __shfl_sync(0xffffffffu, ray.tm, owner);
      Rng rg { __shfl_sync(0xffffffffu, rng.s, owner) };
      const Best b = team_closest_hit<kSmem>(sc, sv, rr, rg, lane % team_size, team_size, active);
      // hand the result back: the j-th live lane reads from the first lane of team j
      const int my_team = __popc(live_mask & ((1u << lane) - 1u));
      const int src = act ? my_team * team_size : 0;
      const float bt = __shfl_sync(0xffffffffu, b.t, src);
      const int bid = __shfl_sync(0xffffffffu, b.id, src);
      const uint32_t brs = __shfl_sync(0xffffffffu, rg.s, src);
      if (act) best.t = bt, best.id = bid, rng.s = brs;
    } else {
      best = closest_hit<kSmem>(sc, sv, ray, rng, act);
    }

    // ---- (C) shade: render.hpp:58-91
    if (act) {
      ++n_scans;
      ++pix_scans;
      V3 contribution = v3(0.f, 0.f, 0.f);
      bool path_done = false;
      if (best.id < 0) {
        // background gradient, render.hpp:83-87
        const V3 ud = unit_vector(ray.d);
        const float hit_pt = fmul(0.5f, fadd(ud.y, 1.0f));
        const float w0 = fsub(1.0f, hit_pt);
        const V3 c = vadd(v3(fmul(w0, 1.0f), fmul(w0, 1.0f), fmul(w0, 1.0f)),
                          v3(fmul(hit_pt, 0.5f), fmul(hit_pt, 0.7f), fmul(hit_pt, 1.0f)));
        contribution = vmul(att, c);
        path_done = true;
      } else {
        HitRec rec;
        const int mat_index = build_record(sc, sv, ray, best, rec, kSmem);
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
        if (scattered_ok) {
          ray = scattered;
          ++bounce;
          if (bounce == p.depth) path_done = true;  // render.hpp:91, black
        } else {
          path_done = true;  // render.hpp:73 (emitted is zero for everything but lights)
        }
      }
      if (path_done) {
        acc = vadd(acc, contribution);
        ++sample;
        need_path = true;
      }
    }
  }

  if (p.counters && (threadIdx.x & 31) == 0) atomicMax(p.counters + 3, globaltimer_ns());  // timeline: warp retired
  // work counters: one atomic per warp
  unsigned int warp_scans = n_scans;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_scans += __shfl_xor_sync(0xffffffffu, warp_scans, o);
  if ((threadIdx.x & 31) == 0 && p.counters) atomicAdd(p.counters, (unsigned long long)warp_scans);
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
  if (info) info->grid = grid, info->block = kBlockThreads, info->smem_bytes = (int)dyn, info->blocks_per_sm = per_sm, info->staged = smem;
  kernel<<<grid, kBlockThreads, dyn, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ptb
