// TEST INFRASTRUCTURE ONLY -- CPU build of the product's closest-hit scan (pt_prims.cuh through
// tests/host/pt_hostshim.h) to check, ray by ray and without a GPU, that every form of culling is INVISIBLE:
// the scan with chunk boxes, flat trees and the grazing index must return the same winner (t bits, object,
// RNG state) as the same scan with all of them switched off, for camera rays, scattered rays and rays
// built to attack the margins (grazing a triangle's plane, aimed at vertices from far away, axis parallel,
// degenerate).  Built and driven by tests/test_flat_culling.py; g++ -O2 -ffp-contract=off -fopenmp.
#include <omp.h>

#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#include "pt_abi.h"
#include "pt_pack.h"
#include "pt_packed.h"
#include "pt_prims.cuh"

using namespace ptb;

namespace {

struct HostScene {
  PackedScene ps;
  std::vector<unsigned char> blob_cull, blob_brute;
  std::vector<int32_t> keys, object_id;
  SceneDesc cull {}, brute {};
};

void fill_desc(const HostScene& h, const std::vector<unsigned char>& blob, SceneDesc& d, uint32_t key_base[6]) {
  const PackedScene& ps = h.ps;
  d.blob = blob.data(), d.blob_bytes = (uint32_t)blob.size(), d.stage_bytes = d.blob_bytes;
  d.n_groups = ps.n_groups;
  d.off_groups = ps.off_groups, d.off_sphere = ps.off_sphere, d.off_moving = ps.off_moving;
  d.off_rect = ps.off_rect, d.off_triangle = ps.off_triangle, d.off_box = ps.off_box;
  d.off_trees = ps.off_trees, d.off_nodes = ps.off_nodes, d.off_tree_ids = ps.off_tree_ids, d.n_trees = ps.n_trees;
  d.off_planes = ps.off_planes, d.n_planes[0] = ps.n_planes[0], d.n_planes[1] = ps.n_planes[1], d.n_planes[2] = ps.n_planes[2];
  d.flat_extent = ps.flat_extent, d.flat_cull = 1u;
  d.n_objects = ps.n_objects;
  d.off_sphere_box = ps.off_sphere_box, d.off_moving_box = ps.off_moving_box;
  d.n_sphere_chunks = (uint32_t)ps.sphere_chunk_open.size(), d.n_moving_chunks = (uint32_t)ps.moving_chunk_open.size();
  d.sphere_aux = ps.sphere_aux.data(), d.moving_aux = ps.moving_aux.data(), d.rect_aux = ps.rect_aux.data();
  d.tri_aux = ps.tri_aux.data(), d.box_aux = ps.box_aux.data(), d.media = ps.media.data();
  d.keys = h.keys.data(), d.object_id = h.object_id.data();
  for (int k = 0; k < 6; ++k) d.key_base[k] = key_base[k];
  d.n_media_groups = ps.n_media_groups, d.n_flat_groups = ps.n_flat_groups, d.n_late_sphere_groups = ps.n_late_sphere_groups;
}

struct Xs {  // xorshift64*
  uint64_t s;
  uint64_t next() {
    s ^= s >> 12, s ^= s << 25, s ^= s >> 27;
    return s * 2685821657736338717ull;
  }
  double u() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double range(double a, double b) { return a + (b - a) * u(); }
  int below(int n) { return (int)(next() % (uint64_t)n); }
};

struct P3 {
  double x, y, z;
};
P3 operator+(P3 a, P3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
P3 operator-(P3 a, P3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
P3 operator*(double s, P3 a) { return { s * a.x, s * a.y, s * a.z }; }
P3 cross(P3 a, P3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
double norm(P3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
P3 random_dir(Xs& g) {
  for (;;) {
    P3 d { g.range(-1, 1), g.range(-1, 1), g.range(-1, 1) };
    const double n = norm(d);
    if (n > 1e-3 && n <= 1) return (1.0 / n) * d;
  }
}

// A random point on (or, with `spread` > 0, around) a random object of the scene.
P3 point_on_object(const pt_scene& sc, Xs& g, double spread, P3* normal, P3* tangent_a, P3* tangent_b) {
  *normal = { 0, 1, 0 }, *tangent_a = { 1, 0, 0 }, *tangent_b = { 0, 0, 1 };
  if (sc.n_hittables == 0) return { 0, 0, 0 };
  const pt_order_entry e = sc.order[g.below((int)sc.n_hittables)];
  const double lo = -spread, hi = 1 + spread;
  switch (e.kind) {
    case PT_HIT_TRIANGLE: {
      const pt_triangle& t = sc.triangles[e.index];
      const P3 v0 { t.v0[0], t.v0[1], t.v0[2] }, e1 { t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2] },
          e2 { t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2] };
      double b = g.range(lo, hi), c = g.range(lo, hi);
      if (spread == 0 && b + c > 1) b = 1 - b, c = 1 - c;
      if (spread == 0 && g.below(4) == 0) b = g.below(2), c = b ? 0 : g.below(2);  // exactly a vertex
      *normal = cross(e1, e2), *tangent_a = e1, *tangent_b = e2;
      return v0 + b * e1 + c * e2;
    }
    case PT_HIT_RECT: {
      const pt_rect& r = sc.rects[e.index];
      const double a = r.a0 + g.range(lo, hi) * (r.a1 - r.a0), b = r.b0 + g.range(lo, hi) * (r.b1 - r.b0);
      if (r.axis == PT_AXIS_XY) {
        *normal = { 0, 0, 1 }, *tangent_a = { 1, 0, 0 }, *tangent_b = { 0, 1, 0 };
        return { a, b, r.k };
      }
      if (r.axis == PT_AXIS_XZ) {
        *normal = { 0, 1, 0 }, *tangent_a = { 1, 0, 0 }, *tangent_b = { 0, 0, 1 };
        return { a, r.k, b };
      }
      *normal = { 1, 0, 0 }, *tangent_a = { 0, 1, 0 }, *tangent_b = { 0, 0, 1 };
      return { r.k, a, b };
    }
    case PT_HIT_BOX: {
      const pt_box& bx = sc.boxes[e.index];
      P3 p { bx.p0[0] + g.range(lo, hi) * (bx.p1[0] - bx.p0[0]), bx.p0[1] + g.range(lo, hi) * (bx.p1[1] - bx.p0[1]),
             bx.p0[2] + g.range(lo, hi) * (bx.p1[2] - bx.p0[2]) };
      const int side = g.below(6);
      double* c = side < 2 ? &p.z : side < 4 ? &p.y : &p.x;
      *c = (side & 1) ? bx.p0[2 - side / 2] : bx.p1[2 - side / 2];
      *normal = side < 2 ? P3 { 0, 0, 1 } : side < 4 ? P3 { 0, 1, 0 } : P3 { 1, 0, 0 };
      *tangent_a = side < 2 ? P3 { 1, 0, 0 } : side < 4 ? P3 { 1, 0, 0 } : P3 { 0, 1, 0 };
      *tangent_b = side < 2 ? P3 { 0, 1, 0 } : side < 4 ? P3 { 0, 0, 1 } : P3 { 0, 0, 1 };
      return p;
    }
    case PT_HIT_SPHERE: {
      const pt_sphere& s = sc.spheres[e.index];
      const P3 n = random_dir(g);
      *normal = n;
      *tangent_a = cross(n, P3 { 0.3, 0.5, 0.8 }), *tangent_b = cross(n, *tangent_a);
      return P3 { s.center0[0], s.center0[1], s.center0[2] } + (std::fabs(s.radius) * (1 + spread * g.range(-1, 1))) * n;
    }
    default: {
      const pt_medium& m = sc.media[e.index];
      if (m.boundary_kind == PT_BOUNDARY_BOX) {
        const pt_box& bx = sc.boxes[m.boundary_index];
        return { bx.p0[0] + g.u() * (bx.p1[0] - bx.p0[0]), bx.p0[1] + g.u() * (bx.p1[1] - bx.p0[1]), bx.p0[2] + g.u() * (bx.p1[2] - bx.p0[2]) };
      }
      const pt_sphere& s = sc.spheres[m.boundary_index];
      return P3 { s.center0[0], s.center0[1], s.center0[2] } + (std::fabs(s.radius) * g.u()) * random_dir(g);
    }
  }
}

double pick(Xs& g, const double* v, int n) { return v[g.below(n)]; }

Ray make_ray(const pt_scene& sc, const pt_camera& cam, int mode, Xs& g, double scene_size) {
  Ray r;
  P3 o, d, n, ta, tb;
  const double tiny[] = { 0, 0, 1e-9, -1e-9, 1e-7, -1e-7, 1e-6, 1e-5, -1e-5, 1e-4, 1e-3, -1e-3, 1e-2 };
  switch (mode) {
    case 0: {  // camera rays (camera.hpp:93-100 without the exact arithmetic: any ray will do)
      const double u = g.u(), v = g.u();
      const P3 org { cam.origin[0], cam.origin[1], cam.origin[2] };
      const P3 off = (cam.lens_radius * g.range(-1, 1)) * P3 { cam.u[0], cam.u[1], cam.u[2] } +
                     (cam.lens_radius * g.range(-1, 1)) * P3 { cam.v[0], cam.v[1], cam.v[2] };
      o = org + off;
      d = P3 { cam.lower_left_corner[0], cam.lower_left_corner[1], cam.lower_left_corner[2] } +
          u * P3 { cam.horizontal[0], cam.horizontal[1], cam.horizontal[2] } + v * P3 { cam.vertical[0], cam.vertical[1], cam.vertical[2] } - org - off;
      break;
    }
    case 1: {  // scattered rays: from a point on an object into a random direction of random length
      o = point_on_object(sc, g, 0.0, &n, &ta, &tb);
      d = g.range(0.05, 2.0) * random_dir(g);
      break;
    }
    case 2: {  // grazing: origin (almost) in an object's plane, direction (almost) inside it
      const P3 p = point_on_object(sc, g, 3.0, &n, &ta, &tb);
      const double nn = norm(n) > 0 ? norm(n) : 1;
      const P3 nu = (1.0 / nn) * n;
      const P3 q = point_on_object(sc, g, 0.0, &n, &ta, &tb);  // (n, ta, tb now belong to another object: mixes planes)
      const bool same_plane = g.below(4) != 0;
      const P3 in_plane = g.range(-1, 1) * ta + g.range(-1, 1) * tb;
      o = p + (pick(g, tiny, 13) * scene_size) * nu;
      d = same_plane ? (q - p) : in_plane;
      if (norm(d) == 0) d = ta;
      d = (g.range(0.2, 1.5) / norm(d)) * d;
      d = d + pick(g, tiny, 13) * nu;
      break;
    }
    case 3: {  // from far away at a point on an object (edges and vertices included), so that |o - v0| is large
      const P3 p = point_on_object(sc, g, 0.0, &n, &ta, &tb);
      const double dist = scene_size * std::pow(10.0, g.range(-1, 4.5));
      const P3 dir = random_dir(g);
      o = p - dist * dir;
      d = g.range(0.3, 3.0) * dir;
      break;
    }
    default: {  // axis parallel, denormal / huge components, NaN and infinite origins
      o = point_on_object(sc, g, 0.5, &n, &ta, &tb) + (scene_size * g.range(0, 0.2)) * random_dir(g);
      d = random_dir(g);
      const double special[] = { 0.0, -0.0, 1e-42, -1e-42, 1e-30, 1e-22, 1e22, -1e25, 1.0 };
      const int which = g.below(8);
      if (which & 1) d.x = pick(g, special, 9);
      if (which & 2) d.y = pick(g, special, 9);
      if (which & 4) d.z = pick(g, special, 9);
      if (which == 0) d = std::pow(10.0, g.range(-30, 30)) * d;
      if (g.below(50) == 0) o.x = std::numeric_limits<double>::quiet_NaN();
      if (g.below(50) == 0) o.y = std::numeric_limits<double>::infinity();
      if (g.below(20) == 0) o = 1e12 * o;
      break;
    }
  }
  r.o = v3((float)o.x, (float)o.y, (float)o.z);
  r.d = v3((float)d.x, (float)d.y, (float)d.z);
  r.tm = (float)g.range(cam.time0, cam.time1);
  return r;
}

}  // namespace

extern "C" {

struct ScanCheckResult {
  uint64_t rays, mismatches, hits;
  uint64_t flat_nodes, graze_nodes, triangle_tests, graze_tests, brute_triangle_tests;
  float bad_ray[7];
  float t_cull, t_brute;
  int id_cull, id_brute;
  int n_trees, tree_levels, tree_leaves, gtree_leaves;
};

// mode < 0: all modes in turn.  Returns 0, or a negative pt error code (message on stderr).
int scan_check(const pt_scene* scene, const pt_camera* cam, int mode, uint64_t seed, uint64_t n_rays, ScanCheckResult* out) {
  HostScene h;
  std::string err;
  const int rc = pack_scene(*scene, h.ps, err);
  if (rc != PT_OK) {
    std::fprintf(stderr, "scan_check: %s\n", err.c_str());
    return rc;
  }
  PackedScene& ps = h.ps;
  uint32_t key_base[6];
  build_key_tables(ps, h.keys, key_base, h.object_id);
  h.keys.push_back(0);

  // brute force: no chunk boxes (the packer's default), no trees
  h.blob_brute = ps.blob;
  for (uint32_t gi = 0; gi < ps.n_groups; ++gi) {
    Group* g = reinterpret_cast<Group*>(h.blob_brute.data() + ps.off_groups) + gi;
    g->tree = -1, g->gtree = -1;
  }
  // culled: the chunk boxes of this camera's shutter interval
  h.blob_cull = ps.blob;
  CullBoxes boxes;
  compute_cull_boxes(ps, cam->time0, cam->time1, boxes);
  if (!boxes.sphere.empty()) std::memcpy(h.blob_cull.data() + ps.off_sphere_box, boxes.sphere.data(), boxes.sphere.size() * sizeof(float));
  if (!boxes.moving.empty()) std::memcpy(h.blob_cull.data() + ps.off_moving_box, boxes.moving.data(), boxes.moving.size() * sizeof(float));
  fill_desc(h, h.blob_cull, h.cull, key_base);
  fill_desc(h, h.blob_brute, h.brute, key_base);
  for (int k = 0; k < 3; ++k) h.cull.cull_bound[k] = boxes.bound[k], h.brute.cull_bound[k] = 0.f;
  const SceneView sv_cull = scene_view(h.cull, h.blob_cull.data()), sv_brute = scene_view(h.brute, h.blob_brute.data());

  double scene_size = 1;
  for (uint32_t i = 0; i < scene->n_triangles; ++i)
    for (int k = 0; k < 3; ++k) scene_size = std::max(scene_size, (double)std::fabs(scene->triangles[i].v0[k]));
  for (uint32_t i = 0; i < scene->n_spheres; ++i)
    if (std::fabs(scene->spheres[i].radius) < 100)
      for (int k = 0; k < 3; ++k) scene_size = std::max(scene_size, (double)std::fabs(scene->spheres[i].center0[k]));

  ScanCheckResult res {};
  res.rays = n_rays;
  res.n_trees = (int)ps.n_trees;
  for (uint32_t gi = 0; gi < ps.n_groups; ++gi) {
    const Group& g = reinterpret_cast<const Group*>(ps.blob.data() + ps.off_groups)[gi];
    const Tree* trees = reinterpret_cast<const Tree*>(ps.blob.data() + ps.off_trees);
    if ((g.type == G_RECT || g.type == G_TRIANGLE || g.type == G_BOX) && g.tree >= 0) {
      res.tree_levels = std::max(res.tree_levels, trees[g.tree].levels), res.tree_leaves += trees[g.tree].n[0];
      if (g.gtree >= 0) res.gtree_leaves += trees[g.gtree].n[0];
    }
  }
  bool have_bad = false;
#pragma omp parallel
  {
    ScanCheckResult loc {};
    pt_host_stats = PtHostStats { 0, 0, 0, 0 };
    unsigned long long brute_tri = 0;
#pragma omp for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n_rays; ++i) {
      Xs g { seed * 0x9e3779b97f4a7c15ull + (uint64_t)i * 0xbf58476d1ce4e5b9ull + 1ull };
      g.next(), g.next();
      const int m = mode >= 0 ? mode : (int)(i % 5);
      const Ray ray = make_ray(*scene, *cam, m, g, scene_size);
      const uint32_t seed0 = (uint32_t)g.next() | 1u;
      Rng rng_a { seed0 }, rng_b { seed0 };
      const Best a = closest_hit<true>(h.cull, sv_cull, ray, rng_a, true, 0, 1);
      const unsigned long long tri_before = pt_host_stats.triangle_tests;
      const PtHostStats keep = pt_host_stats;
      const Best b = closest_hit<true>(h.brute, sv_brute, ray, rng_b, true, 0, 1);
      brute_tri += pt_host_stats.triangle_tests - tri_before;
      pt_host_stats = keep;
      // ... and the reference's own scan, object by object in vector order: the (min t, max key) rule over groups re-ordered
      // by kind must give the same winner whenever no NaN is in play
      Rng rng_c { seed0 };
      const Best c = closest_hit_in_order<true>(h.brute, sv_brute, ray, rng_c);
      const bool nan_in_play = c.t != c.t || b.t != b.t;
      const bool order_ok = nan_in_play || (c.id == b.id && (c.id < 0 || __float_as_uint(c.t) == __float_as_uint(b.t)) && rng_c.s == rng_b.s);
      const bool same = order_ok && a.id == b.id && (a.id < 0 || __float_as_uint(a.t) == __float_as_uint(b.t)) && rng_a.s == rng_b.s;
      if (a.id >= 0) ++loc.hits;
      if (!same) {
        ++loc.mismatches;
#pragma omp critical
        if (!have_bad) {
          have_bad = true;
          const float rr[7] = { ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, ray.tm };
          std::memcpy(res.bad_ray, rr, sizeof rr);
          res.t_cull = order_ok ? a.t : c.t, res.t_brute = b.t, res.id_cull = order_ok ? a.id : c.id, res.id_brute = b.id;
        }
      }
    }
#pragma omp critical
    {
      res.mismatches += loc.mismatches, res.hits += loc.hits;
      res.flat_nodes += pt_host_stats.flat_nodes, res.graze_nodes += pt_host_stats.graze_nodes;
      res.triangle_tests += pt_host_stats.triangle_tests, res.graze_tests += pt_host_stats.graze_tests;
      res.brute_triangle_tests += brute_tri;
    }
  }
  *out = res;
  return 0;
}

}  // extern "C"

// The rays of scan_check() (same generator, same seeds) for callers that trace them elsewhere -- the GPU ray-level
// parity test feeds them to the CUDA scan and to the oracle's hit_world.
extern "C" int scan_check_make_rays(const pt_scene* scene, const pt_camera* cam, int mode, uint64_t seed, uint64_t n_rays, float* rays7,
                                    uint32_t* seeds) {
  double scene_size = 1;
  for (uint32_t i = 0; i < scene->n_triangles; ++i)
    for (int k = 0; k < 3; ++k) scene_size = std::max(scene_size, (double)std::fabs(scene->triangles[i].v0[k]));
  for (uint32_t i = 0; i < scene->n_spheres; ++i)
    if (std::fabs(scene->spheres[i].radius) < 100)
      for (int k = 0; k < 3; ++k) scene_size = std::max(scene_size, (double)std::fabs(scene->spheres[i].center0[k]));
  for (uint64_t i = 0; i < n_rays; ++i) {
    Xs g { seed * 0x9e3779b97f4a7c15ull + i * 0xbf58476d1ce4e5b9ull + 1ull };
    g.next(), g.next();
    const int m = mode >= 0 ? mode : (int)(i % 5);
    const Ray ray = make_ray(*scene, *cam, m, g, scene_size);
    const float rr[7] = { ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, ray.tm };
    std::memcpy(rays7 + 7 * i, rr, sizeof rr);
    seeds[i] = (uint32_t)g.next() | 1u;
  }
  return 0;
}
