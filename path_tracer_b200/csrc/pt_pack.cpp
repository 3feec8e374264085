// pt_pack.cpp -- pt_scene -> device layout (see pt_packed.h).
//
// Everything the reference derives from the constructor arguments at scene
// build time or re-derives per hit() call from loop-invariant data is derived
// here ONCE with the same single IEEE-754 binary32 operation, so the device
// sees bit-identical values:
//   radius*radius            sphere.hpp:71
//   center1 - center0        sphere.hpp:55
//   time1 - time0            sphere.hpp:55
//   v1 - v0, v2 - v0         triangle.hpp:65-66
//   cross(edge1, edge2)      triangle.hpp:96
//   -1 / density             constant_medium.hpp:20
#include "pt_pack.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <utility>

namespace ptb {
namespace {

struct f4 {
  float x, y, z, w;
};

struct RawSphere {
  f4 head;  // {c0, r*r inflated for the miss filter}
  f4 dv;    // {c1 - c0, 0}: moving spheres only
  SphereAux aux;
  SphereGeo geo;
  bool outsized;  // never culled, kept apart from the k-d ordered chunks
};

struct MovingClass {
  float time0, time1;
  std::vector<RawSphere> items;
};

struct Segment {
  std::vector<RawSphere> sph;
  std::vector<MovingClass> classes;
  std::vector<f4> rect;
  std::vector<ObjAux> rect_aux;
  std::vector<f4> tri;
  std::vector<TriAux> tri_aux;
  std::vector<f4> box;
  std::vector<ObjAux> box_aux;
  bool empty() const { return sph.empty() && classes.empty() && rect.empty() && tri.empty() && box.empty(); }
};

float mid_coord(const RawSphere& s, int axis) { return 0.5f * s.geo.c0[axis] + 0.5f * s.geo.c1[axis]; }

// k-d order: split at the multiple of kSphereChunk nearest to the median along the widest axis of the
// centres, so that every run of kSphereChunk spheres is a compact cluster (all chunks but one are full).
void kd_order(std::vector<RawSphere>& v, size_t lo, size_t hi) {
  const size_t n = hi - lo;
  if (n <= (size_t)kSphereChunk) return;
  int axis = 0;
  float widest = -1.f;
  for (int k = 0; k < 3; ++k) {
    float mn = std::numeric_limits<float>::infinity(), mx = -mn;
    for (size_t i = lo; i < hi; ++i) mn = std::min(mn, mid_coord(v[i], k)), mx = std::max(mx, mid_coord(v[i], k));
    if (mx - mn > widest) widest = mx - mn, axis = k;
  }
  const size_t mid = lo + ((n / 2 + kSphereChunk - 1) / kSphereChunk) * kSphereChunk;
  std::nth_element(v.begin() + (long)lo, v.begin() + (long)mid, v.begin() + (long)hi,
                   [axis](const RawSphere& a, const RawSphere& b) { return mid_coord(a, axis) < mid_coord(b, axis); });
  kd_order(v, lo, mid);
  kd_order(v, mid, hi);
}


// ---------------------------------------------------------------- flat groups: k-d order, box trees, grazing index
struct DBox {  // bounding box in binary64; `open` = not finite: never culled
  double lo[3], hi[3];
  bool open;
};
DBox box_of_points(const double (*p)[3], int n) {
  DBox b;
  b.open = false;
  for (int k = 0; k < 3; ++k) {
    b.lo[k] = std::numeric_limits<double>::infinity(), b.hi[k] = -b.lo[k];
    for (int i = 0; i < n; ++i) {
      if (!std::isfinite(p[i][k])) b.open = true;
      b.lo[k] = std::min(b.lo[k], p[i][k]), b.hi[k] = std::max(b.hi[k], p[i][k]);
    }
  }
  return b;
}
void grow(DBox& a, const DBox& b) {
  a.open = a.open || b.open;
  for (int k = 0; k < 3; ++k) a.lo[k] = std::min(a.lo[k], b.lo[k]), a.hi[k] = std::max(a.hi[k], b.hi[k]);
}
float round_down(double x) {
  float f = (float)x;
  if ((double)f > x) f = std::nextafter(f, -std::numeric_limits<float>::infinity());
  return f;
}
float round_up(double x) {
  float f = (float)x;
  if ((double)f < x) f = std::nextafter(f, std::numeric_limits<float>::infinity());
  return f;
}

// k-d order of perm[lo, hi) by the points `pts`: median splits along the widest axis, at multiples of the largest
// leaf * kTreeFan^k below the range's size, so that EVERY aligned run of leaf * kTreeFan^k elements (= a tree
// node) is a union of k-d cells, i.e. spatially compact.
void kd_order_aligned(std::vector<int>& perm, const std::vector<std::array<double, 3>>& pts, size_t lo, size_t hi, size_t leaf) {
  const size_t n = hi - lo;
  if (n <= leaf) return;
  size_t block = leaf;
  while (block * (size_t)kTreeFan < n) block *= (size_t)kTreeFan;
  int axis = 0;
  double widest = -1;
  for (int k = 0; k < 3; ++k) {
    double mn = std::numeric_limits<double>::infinity(), mx = -mn;
    for (size_t i = lo; i < hi; ++i) mn = std::min(mn, pts[(size_t)perm[i]][(size_t)k]), mx = std::max(mx, pts[(size_t)perm[i]][(size_t)k]);
    if (mx - mn > widest) widest = mx - mn, axis = k;
  }
  size_t k = (n / 2 + block / 2) / block;
  if (k < 1) k = 1;
  if (k * block >= n) k = (n - 1) / block;
  const size_t mid = lo + k * block;
  std::nth_element(perm.begin() + (long)lo, perm.begin() + (long)mid, perm.begin() + (long)hi,
                   [&pts, axis](int a, int b) { return pts[(size_t)a][(size_t)axis] < pts[(size_t)b][(size_t)axis]; });
  kd_order_aligned(perm, pts, lo, mid, leaf);
  kd_order_aligned(perm, pts, mid, hi, leaf);
}

// Box tree over elements that are already in their final order: leaves of `leaf` consecutive elements, kTreeFan
// children per inner node, boxes rounded outward to binary32 ({lo} {hi} float4 pairs appended to `nodes`).
Tree build_tree(const std::vector<DBox>& elem, size_t leaf, std::vector<f4>& nodes, double* extent) {
  Tree t {};
  t.leaf_ids = -1;
  std::vector<DBox> level;
  for (size_t i = 0; i < elem.size(); i += leaf) {
    DBox b = elem[i];
    for (size_t j = i + 1; j < std::min(elem.size(), i + leaf); ++j) grow(b, elem[j]);
    level.push_back(b);
  }
  const float inf = std::numeric_limits<float>::infinity();
  for (int l = 0; l < kTreeLevels; ++l) {
    t.off[l] = (int32_t)nodes.size(), t.n[l] = (int32_t)level.size(), t.levels = l + 1;
    for (const DBox& b : level) {
      if (b.open) {
        nodes.push_back(f4 { -inf, -inf, -inf, 0.f }), nodes.push_back(f4 { inf, inf, inf, 0.f });
      } else {
        nodes.push_back(f4 { round_down(b.lo[0]), round_down(b.lo[1]), round_down(b.lo[2]), 0.f });
        nodes.push_back(f4 { round_up(b.hi[0]), round_up(b.hi[1]), round_up(b.hi[2]), 0.f });
        if (extent)
          for (int k = 0; k < 3; ++k) *extent = std::max(*extent, std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
      }
    }
    if (level.size() <= 32 || l + 1 == kTreeLevels) break;
    std::vector<DBox> up;
    for (size_t i = 0; i < level.size(); i += (size_t)kTreeFan) {
      DBox b = level[i];
      for (size_t j = i + 1; j < std::min(level.size(), i + (size_t)kTreeFan); ++j) grow(b, level[j]);
      up.push_back(b);
    }
    level.swap(up);
  }
  return t;
}

// Reorder the `per` float4 per element of `data` (and the matching aux entries) by `perm`.
template <typename Aux> void permute_flat(std::vector<f4>& data, std::vector<Aux>& aux, const std::vector<int>& perm, int per) {
  std::vector<f4> d2(data.size());
  std::vector<Aux> a2(aux.size());
  for (size_t i = 0; i < perm.size(); ++i) {
    for (int j = 0; j < per; ++j) d2[i * (size_t)per + (size_t)j] = data[(size_t)perm[i] * (size_t)per + (size_t)j];
    a2[i] = aux[(size_t)perm[i]];
  }
  data.swap(d2), aux.swap(a2);
}

// Bounding box of one flat element (binary64 from the binary32 data the device sees).
DBox flat_box(int type, const f4* e) {
  double p[3][3];
  if (type == G_TRIANGLE) {  // v0, v0 + e1, v0 + e2 with the hoisted edges (triangle.hpp:65-66)
    for (int k = 0; k < 3; ++k) {
      const double v0 = (&e[0].x)[k];
      p[0][k] = v0, p[1][k] = v0 + (double)(&e[1].x)[k], p[2][k] = v0 + (double)(&e[2].x)[k];
    }
    return box_of_points(p, 3);
  }
  if (type == G_RECT) {  // {a0, a1, b0, b1} {k, axis}
    int32_t axis;
    std::memcpy(&axis, &e[1].y, 4);
    const int ia = axis == PT_AXIS_YZ ? 1 : 0, ib = axis == PT_AXIS_XY ? 1 : 2, ik = axis == PT_AXIS_XY ? 2 : axis == PT_AXIS_XZ ? 1 : 0;
    p[0][ia] = e[0].x, p[1][ia] = e[0].y, p[0][ib] = e[0].z, p[1][ib] = e[0].w, p[0][ik] = p[1][ik] = e[1].x;
    return box_of_points(p, 2);
  }
  for (int k = 0; k < 3; ++k) p[0][k] = (&e[0].x)[k], p[1][k] = (&e[1].x)[k];  // box: p0, p1
  return box_of_points(p, 2);
}

struct Builder {
  std::vector<Group> groups;
  std::vector<f4> sph, mov, rect, tri, box;
  std::vector<Tree> trees;
  std::vector<f4> nodes;
  std::vector<f4> tree_ids;  // grazing index leaves: {g, element index as bits} per triangle
  double flat_extent = 0;
  PackedScene& out;
  explicit Builder(PackedScene& o) : out(o) {}

  static RawSphere padding() {
    // A padding sphere can never become a candidate: r*r = -inf makes
    // c = dot(oc,oc) - r*r = +inf, so discriminant = b*b - a*c is -inf or NaN.
    RawSphere p {};
    p.head = f4 { 0.f, 0.f, 0.f, -std::numeric_limits<float>::infinity() };
    p.dv = f4 { 0.f, 0.f, 0.f, 0.f };
    p.aux.radius = 1.f, p.aux.material = -1, p.aux.key = std::numeric_limits<int32_t>::min();
    p.geo.valid = false;
    return p;
  }

  // Chunk layout of pt_packed.h: outsized spheres first (chunks that are never culled), then the k-d
  // ordered rest; every chunk's entries are written twice in a row.
  static int emit_spheres(std::vector<RawSphere> items, bool moving, std::vector<f4>& data,
                          std::vector<SphereAux>& aux, std::vector<SphereGeo>& geo, std::vector<unsigned char>& open,
                          int& n_outsized) {
    std::vector<RawSphere> ordered;
    size_t n_open_chunks = 0;
    for (const RawSphere& s : items)
      if (s.outsized) ordered.push_back(s);
    n_outsized = (int)ordered.size();
    while (ordered.size() % kSphereChunk) ordered.push_back(padding());
    n_open_chunks = ordered.size() / kSphereChunk;
    std::vector<RawSphere> rest;
    for (const RawSphere& s : items)
      if (!s.outsized) rest.push_back(s);
    kd_order(rest, 0, rest.size());
    ordered.insert(ordered.end(), rest.begin(), rest.end());
    while (ordered.size() % kSphereChunk) ordered.push_back(padding());
    for (size_t c = 0; c < ordered.size() / kSphereChunk; ++c) {
      const RawSphere* ch = ordered.data() + c * kSphereChunk;
      for (int rep = 0; rep < 2; ++rep)
        for (int k = 0; k < kSphereChunk; ++k) data.push_back(ch[k].head);
      if (moving)
        for (int rep = 0; rep < 2; ++rep)
          for (int k = 0; k < kSphereChunk; ++k) data.push_back(ch[k].dv);
      for (int k = 0; k < kSphereChunk; ++k) aux.push_back(ch[k].aux), geo.push_back(ch[k].geo);
      open.push_back(c < n_open_chunks ? 1 : 0);
    }
    return (int)ordered.size();
  }

  void flush(Segment& s) {
    if (!s.sph.empty()) {
      Group g {};
      g.type = G_SPHERE, g.begin = (int32_t)out.sphere_aux.size(), g.tree = g.gtree = -1;
      g.count = emit_spheres(s.sph, false, sph, out.sphere_aux, out.sphere_geo, out.sphere_chunk_open, g.n_open);
      groups.push_back(g);
    }
    for (auto& c : s.classes) {
      Group g {};
      g.type = G_MOVING_SPHERE, g.begin = (int32_t)out.moving_aux.size(), g.tree = g.gtree = -1;
      g.count = emit_spheres(c.items, true, mov, out.moving_aux, out.moving_geo, out.moving_chunk_open, g.n_open);
      g.time0 = c.time0, g.den = c.time1 - c.time0;
      groups.push_back(g);
    }
    if (!s.rect_aux.empty()) {
      Group g {};
      g.type = G_RECT, g.begin = (int32_t)out.rect_aux.size(), g.count = (int32_t)s.rect_aux.size();
      flat_tree(g, s.rect, s.rect_aux, 2);
      groups.push_back(g);
      rect.insert(rect.end(), s.rect.begin(), s.rect.end());
      out.rect_aux.insert(out.rect_aux.end(), s.rect_aux.begin(), s.rect_aux.end());
    }
    if (!s.tri_aux.empty()) {
      Group g {};
      g.type = G_TRIANGLE, g.begin = (int32_t)out.tri_aux.size(), g.count = (int32_t)s.tri_aux.size();
      flat_tree(g, s.tri, s.tri_aux, 3);
      groups.push_back(g);
      tri.insert(tri.end(), s.tri.begin(), s.tri.end());
      out.tri_aux.insert(out.tri_aux.end(), s.tri_aux.begin(), s.tri_aux.end());
    }
    if (!s.box_aux.empty()) {
      Group g {};
      g.type = G_BOX, g.begin = (int32_t)out.box_aux.size(), g.count = (int32_t)s.box_aux.size();
      flat_tree(g, s.box, s.box_aux, 2);
      groups.push_back(g);
      box.insert(box.end(), s.box.begin(), s.box.end());
      out.box_aux.insert(out.box_aux.end(), s.box_aux.begin(), s.box_aux.end());
    }
    s = Segment {};
  }

  // A flat group worth culling: k-d order its elements (the winner rule is order independent, pt_packed.h), hang the
  // leaves of kFlatChunk elements under a box tree and, for triangles, build the grazing index.
  template <typename Aux> void flat_tree(Group& g, std::vector<f4>& data, std::vector<Aux>& aux, int per) {
    g.tree = -1, g.gtree = -1;
    const size_t n = aux.size();
    if (n < (size_t)kFlatTreeMin) return;
    std::vector<DBox> boxes(n);
    std::vector<std::array<double, 3>> mid(n);
    for (size_t i = 0; i < n; ++i) {
      boxes[i] = flat_box(g.type, data.data() + i * (size_t)per);
      for (size_t k = 0; k < 3; ++k) mid[i][k] = boxes[i].open ? 0.0 : 0.5 * boxes[i].lo[k] + 0.5 * boxes[i].hi[k];
    }
    std::vector<int> perm(n);
    for (size_t i = 0; i < n; ++i) perm[i] = (int)i;
    kd_order_aligned(perm, mid, 0, n, (size_t)kFlatChunk);
    permute_flat(data, aux, perm, per);
    std::vector<DBox> ordered(n);
    for (size_t i = 0; i < n; ++i) ordered[i] = boxes[(size_t)perm[i]];
    g.tree = (int32_t)trees.size();
    trees.push_back(build_tree(ordered, (size_t)kFlatChunk, nodes, &flat_extent));
    if (g.type != G_TRIANGLE) return;
    // GRAZING INDEX: g = cross(e1, e2) / (|e1| |e2|) per triangle, sign-normalised (only |d . g| matters); a triangle
    // with a zero or non-finite edge has a = 0 or lives in an open leaf (always visited) and needs no entry.
    std::vector<int> members;
    std::vector<std::array<double, 3>> gv(n);
    for (size_t i = 0; i < n; ++i) {
      const f4* e = data.data() + i * 3;
      const double e1[3] = { e[1].x, e[1].y, e[1].z }, e2[3] = { e[2].x, e[2].y, e[2].z };
      const double l1 = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]), l2 = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
      const double den = l1 * l2;
      gv[i] = { 0, 0, 0 };
      if (!(den > 0) || !std::isfinite(den)) continue;
      double c[3] = { (e1[1] * e2[2] - e1[2] * e2[1]) / den, (e1[2] * e2[0] - e1[0] * e2[2]) / den, (e1[0] * e2[1] - e1[1] * e2[0]) / den };
      int big = 0;
      for (int k = 1; k < 3; ++k)
        if (std::fabs(c[k]) > std::fabs(c[big])) big = k;
      if (c[big] < 0) c[0] = -c[0], c[1] = -c[1], c[2] = -c[2];
      gv[i] = { c[0], c[1], c[2] };
      members.push_back((int)i);
    }
    if (members.empty()) return;
    kd_order_aligned(members, gv, 0, members.size(), (size_t)kFlatChunk);
    std::vector<DBox> gboxes(members.size());
    for (size_t i = 0; i < members.size(); ++i) {
      const double p[1][3] = { { gv[(size_t)members[i]][0], gv[(size_t)members[i]][1], gv[(size_t)members[i]][2] } };
      gboxes[i] = box_of_points(p, 1);
    }
    Tree gt = build_tree(gboxes, (size_t)kFlatChunk, nodes, nullptr);
    gt.leaf_ids = (int32_t)tree_ids.size();
    for (size_t i = 0; i < members.size(); ++i) {
      const int32_t id = g.begin + members[i];
      f4 e { (float)gv[(size_t)members[i]][0], (float)gv[(size_t)members[i]][1], (float)gv[(size_t)members[i]][2], 0.f };
      std::memcpy(&e.w, &id, 4);
      tree_ids.push_back(e);
    }
    while (tree_ids.size() % (size_t)kFlatChunk) {
      f4 e { 0.f, 0.f, 0.f, 0.f };
      const int32_t none = -1;
      std::memcpy(&e.w, &none, 4);
      tree_ids.push_back(e);
    }
    g.gtree = (int32_t)trees.size();
    trees.push_back(gt);
  }
};

// |centre| + |radius|, the largest over the sphere's own motion
double sphere_extent(const pt_sphere& s) {
  double e = 0;
  for (int end = 0; end < 2; ++end) {
    const float* c = end ? s.center1 : s.center0;
    e = std::max(e, std::sqrt((double)c[0] * c[0] + (double)c[1] * c[1] + (double)c[2] * c[2]));
  }
  return e + std::fabs((double)s.radius);
}

uint32_t append(std::vector<unsigned char>& blob, const void* p, size_t bytes) {
  while (blob.size() % 16) blob.push_back(0);
  const uint32_t off = (uint32_t)blob.size();
  const unsigned char* b = static_cast<const unsigned char*>(p);
  blob.insert(blob.end(), b, b + bytes);
  return off;
}

}  // namespace

// CHUNK BOXES.  A sphere the reference hits at parameter t puts the point o + t d within
// sqrt(r^2 + E) of the (float) centre, E = 48 u (|o - c| + r)^2: the residual of the float root in the
// exact quadratic (error analysis in DESIGN.md, "chunk culling").  A box set serves every origin with
// max |coordinate| <= bound, so |o - c| + r <= sqrt(3) bound + S with S = max (|c| + |r|) over the culled
// spheres, and each chunk's box is the union of its spheres' boxes (over the sweep of the moving
// ones) grown by  m = 2 min(sqrt(E), E / (2 r_min)) + 1e-6 (bound + S) (2 + |f|max);  the second term
// covers the rounding of the centre, of the slab test itself and of the ray time.
void compute_cull_boxes(const PackedScene& ps, float cam_time0, float cam_time1, CullBoxes& out) {
  const double inf = std::numeric_limits<double>::infinity();
  struct Range {
    double lo[3], hi[3];
    bool any;
  };
  const std::vector<SphereGeo>* geos[2] = { &ps.sphere_geo, &ps.moving_geo };
  const std::vector<unsigned char>* opens[2] = { &ps.sphere_chunk_open, &ps.moving_chunk_open };
  std::vector<Range> ranges[2];
  double s_max = 0, r_min = inf, f_abs_max = 0;
  bool usable = cam_time0 == cam_time0 && cam_time1 == cam_time1 && std::isfinite(cam_time0) && std::isfinite(cam_time1);
  const double ct_lo = std::min(cam_time0, cam_time1), ct_hi = std::max(cam_time0, cam_time1);
  const double ct_pad = 1e-6 * (std::fabs(ct_lo) + std::fabs(ct_hi));
  for (int kind = 0; kind < 2; ++kind) {
    const size_t n_chunks = geos[kind]->size() / kSphereChunk;
    ranges[kind].assign(n_chunks, Range { { inf, inf, inf }, { -inf, -inf, -inf }, false });
    for (size_t i = 0; i < geos[kind]->size(); ++i) {
      const SphereGeo& g = (*geos[kind])[i];
      if (!g.valid) continue;
      Range& rg = ranges[kind][i / kSphereChunk];
      rg.any = true;
      if ((*opens[kind])[i / kSphereChunk]) continue;
      double f[2] = { 0, 0 };
      if (g.time0 != g.time1) {
        const double den = (double)(g.time1 - g.time0);  // the float the device divides by (sphere.hpp:55)
        const double fa = (ct_lo - ct_pad - g.time0) / den, fb = (ct_hi + ct_pad - g.time0) / den;
        const double pad = 1e-5 * (1.0 + std::max(std::fabs(fa), std::fabs(fb)));
        f[0] = std::min(fa, fb) - pad, f[1] = std::max(fa, fb) + pad;
        if (!(std::isfinite(f[0]) && std::isfinite(f[1]))) usable = false;
        f_abs_max = std::max(f_abs_max, std::max(std::fabs(f[0]), std::fabs(f[1])));
      }
      const double r = std::fabs((double)g.radius);
      r_min = std::min(r_min, r);
      for (int end = 0; end < 2; ++end) {
        double c[3], len2 = 0;
        for (int k = 0; k < 3; ++k) {
          const double dv = (double)(float)(g.c1[k] - g.c0[k]);  // the float the device multiplies by
          c[k] = g.c0[k] + f[end] * dv;
          len2 += c[k] * c[k];
          rg.lo[k] = std::min(rg.lo[k], c[k] - r), rg.hi[k] = std::max(rg.hi[k], c[k] + r);
        }
        s_max = std::max(s_max, std::sqrt(len2) + r);
      }
    }
  }
  if (!(s_max > 0 && s_max < 1e12)) usable = false;
  for (int k = 0; k < 3; ++k) out.bound[k] = 0.f;
  double margin[kCullSets] = { inf, inf, inf, inf };
  if (usable) {
    const double mult[3] = { 2, 16, 128 };
    for (int s = 0; s < 3; ++s) {
      const double bound = mult[s] * s_max;
      out.bound[s] = std::nextafter((float)bound, 0.f);
      const double d = 1.7321 * bound + s_max;
      const double e = 4e-6 * d * d;
      const double m_geom = std::min(std::sqrt(e), r_min > 0 ? e / (2 * r_min) : inf);
      margin[s] = 2 * m_geom + 1e-6 * (bound + s_max) * (2 + f_abs_max);
    }
  }
  for (int kind = 0; kind < 2; ++kind) {
    std::vector<float>& dst = kind ? out.moving : out.sphere;
    const size_t n_chunks = ranges[kind].size();
    dst.assign((size_t)kCullSets * n_chunks * 8, 0.f);
    for (int s = 0; s < kCullSets; ++s)
      for (size_t c = 0; c < n_chunks; ++c) {
        float* b = dst.data() + ((size_t)s * n_chunks + c) * 8;
        const Range& rg = ranges[kind][c];
        const float finf = std::numeric_limits<float>::infinity();
        bool open = (*opens[kind])[c] || !(margin[s] < inf);
        for (int k = 0; k < 3 && !open; ++k)
          if (!(std::fabs(rg.lo[k] - margin[s]) < 1e15 && std::fabs(rg.hi[k] + margin[s]) < 1e15)) open = true;
        for (int k = 0; k < 3; ++k) {
          if (!rg.any) {
            b[k] = finf, b[4 + k] = -finf;  // nothing but padding: never scanned
          } else if (open) {
            b[k] = -finf, b[4 + k] = finf;
          } else {
            b[k] = std::nextafter((float)(rg.lo[k] - margin[s]), -finf);
            b[4 + k] = std::nextafter((float)(rg.hi[k] + margin[s]), finf);
          }
        }
      }
  }
}

void build_key_tables(const PackedScene& ps, std::vector<int32_t>& keys, uint32_t key_base[6], std::vector<int32_t>& object_id) {
  keys.clear();
  key_base[G_SPHERE] = (uint32_t)keys.size();
  for (const auto& a : ps.sphere_aux) keys.push_back(a.key);
  key_base[G_MOVING_SPHERE] = (uint32_t)keys.size();
  for (const auto& a : ps.moving_aux) keys.push_back(a.key);
  key_base[G_RECT] = (uint32_t)keys.size();
  for (const auto& a : ps.rect_aux) keys.push_back(a.key);
  key_base[G_TRIANGLE] = (uint32_t)keys.size();
  for (const auto& a : ps.tri_aux) keys.push_back(a.key);
  key_base[G_BOX] = (uint32_t)keys.size();
  for (const auto& a : ps.box_aux) keys.push_back(a.key);
  key_base[G_MEDIUM] = (uint32_t)keys.size();
  for (const auto& a : ps.media) keys.push_back(a.key);
  // a sphere's key is -1 - index, any other object's is the index (pt_packed.h); padding spheres have no material
  object_id.assign(std::max<uint32_t>(ps.n_objects, 1u), -1);
  for (size_t i = 0; i < ps.rect_aux.size(); ++i) object_id[(size_t)ps.rect_aux[i].key] = make_id(G_RECT, (int)i);
  for (size_t i = 0; i < ps.tri_aux.size(); ++i) object_id[(size_t)ps.tri_aux[i].key] = make_id(G_TRIANGLE, (int)i);
  for (size_t i = 0; i < ps.box_aux.size(); ++i) object_id[(size_t)ps.box_aux[i].key] = make_id(G_BOX, (int)i);
  for (size_t i = 0; i < ps.media.size(); ++i) object_id[(size_t)ps.media[i].key] = make_id(G_MEDIUM, (int)i);
  for (size_t i = 0; i < ps.sphere_aux.size(); ++i)
    if (ps.sphere_aux[i].material >= 0) object_id[(size_t)(-1 - ps.sphere_aux[i].key)] = make_id(G_SPHERE, (int)i);
  for (size_t i = 0; i < ps.moving_aux.size(); ++i)
    if (ps.moving_aux[i].material >= 0) object_id[(size_t)(-1 - ps.moving_aux[i].key)] = make_id(G_MOVING_SPHERE, (int)i);
}

int pack_scene(const pt_scene& sc, PackedScene& out, std::string& error) {
  out = PackedScene {};
  if ((sc.n_hittables && !sc.order) || (sc.n_spheres && !sc.spheres) || (sc.n_rects && !sc.rects) ||
      (sc.n_triangles && !sc.triangles) || (sc.n_boxes && !sc.boxes) || (sc.n_media && !sc.media) ||
      (sc.n_materials && !sc.materials) || (sc.n_textures && !sc.textures) ||
      (sc.n_texture_bytes && !sc.texture_bytes)) {
    error = "pt_scene: null array with non-zero count";
    return PT_ERR_INVALID_ARGUMENT;
  }
  // materials / textures are used verbatim on the device; validate references
  for (uint32_t i = 0; i < sc.n_textures; ++i) {
    const pt_texture& t = sc.textures[i];
    if (t.kind < PT_TEX_CHECKER || t.kind > PT_TEX_IMAGE) {
      error = "pt_scene: unknown texture kind";
      return PT_ERR_INVALID_ARGUMENT;
    }
    if (t.kind == PT_TEX_IMAGE) {
      if (t.width == 0 || t.height == 0 ||
          (t.offset + (uint64_t)t.width * t.height) * 3u > sc.n_texture_bytes) {
        error = "pt_scene: image texture outside the texture byte pool";
        return PT_ERR_INVALID_ARGUMENT;
      }
    }
  }
  for (uint32_t i = 0; i < sc.n_materials; ++i) {
    const pt_material& m = sc.materials[i];
    if (m.kind < PT_MAT_LAMBERTIAN || m.kind > PT_MAT_ISOTROPIC) {
      error = "pt_scene: unknown material kind";
      return PT_ERR_INVALID_ARGUMENT;
    }
    const bool textured = m.kind == PT_MAT_LAMBERTIAN || m.kind == PT_MAT_LIGHTSOURCE || m.kind == PT_MAT_ISOTROPIC;
    if (textured && (m.texture < 0 || (uint32_t)m.texture >= sc.n_textures)) {
      error = "pt_scene: material references a texture out of range";
      return PT_ERR_INVALID_ARGUMENT;
    }
  }
  auto mat_ok = [&](int32_t m) { return m >= 0 && (uint32_t)m < sc.n_materials; };

  // Spheres far larger than the rest (a ground sphere of radius 1000 among marbles) would blow up the
  // culling margins, which grow with the extent of what is culled: they go to chunks of their own
  // that are always scanned.  "Far larger" = more than 4 x the median of |centre| + |radius|.
  double outsized_above = std::numeric_limits<double>::infinity();
  {
    std::vector<double> ext;
    for (uint32_t i = 0; i < sc.n_hittables; ++i)
      if (sc.order[i].kind == PT_HIT_SPHERE && sc.order[i].index >= 0 && (uint32_t)sc.order[i].index < sc.n_spheres)
        ext.push_back(sphere_extent(sc.spheres[sc.order[i].index]));
    if (!ext.empty()) {
      std::nth_element(ext.begin(), ext.begin() + (long)((ext.size() - 1) / 2), ext.end());
      outsized_above = 4.0 * ext[(ext.size() - 1) / 2];
    }
  }

  Builder b(out);
  Segment seg;
  std::vector<float> planes[3];  // plane coordinates of rectangles, box sides and media boundary boxes, per component
  for (uint32_t i = 0; i < sc.n_hittables; ++i) {
    const pt_order_entry e = sc.order[i];
    switch (e.kind) {
      case PT_HIT_SPHERE: {
        if (e.index < 0 || (uint32_t)e.index >= sc.n_spheres || !mat_ok(sc.spheres[e.index].material)) {
          error = "pt_scene: bad sphere reference";
          return PT_ERR_INVALID_ARGUMENT;
        }
        const pt_sphere& s = sc.spheres[e.index];
        RawSphere rs {};
        SphereAux& a = rs.aux;
        a.radius = s.radius, a.material = s.material, a.key = -1 - (int32_t)i;
        // The scan blob carries r*r inflated by the miss filter's margin, rounded up (pt_kernel.cu,
        // "conservative miss filter"); the exact r*r is recomputed from the side table's radius.
        const float r2_exact = s.radius * s.radius;
        const float r2 = std::nextafter(r2_exact * (1.0f + 2.0f * 4.0e-6f / (1.0f - 4.0e-6f)),
                                        std::numeric_limits<float>::infinity());
        rs.head = f4 { s.center0[0], s.center0[1], s.center0[2], r2 };
        rs.geo.radius = s.radius, rs.geo.time0 = s.time0, rs.geo.time1 = s.time1, rs.geo.valid = true;
        for (int k = 0; k < 3; ++k) rs.geo.c0[k] = s.center0[k], rs.geo.c1[k] = s.center0[k];
        rs.outsized = !(sphere_extent(s) <= outsized_above);
        if (s.time0 == s.time1) {  // sphere.hpp:52
          seg.sph.push_back(rs);
        } else {
          MovingClass* cls = nullptr;
          for (auto& c : seg.classes)
            if (c.time0 == s.time0 && c.time1 == s.time1) cls = &c;
          if (!cls) {
            seg.classes.push_back(MovingClass { s.time0, s.time1, {} });
            cls = &seg.classes.back();
          }
          a.time0 = s.time0, a.den = s.time1 - s.time0;
          rs.dv = f4 { s.center1[0] - s.center0[0], s.center1[1] - s.center0[1], s.center1[2] - s.center0[2], 0.f };
          for (int k = 0; k < 3; ++k) rs.geo.c1[k] = s.center1[k];
          cls->items.push_back(rs);
        }
        break;
      }
      case PT_HIT_RECT: {
        if (e.index < 0 || (uint32_t)e.index >= sc.n_rects || !mat_ok(sc.rects[e.index].material) ||
            sc.rects[e.index].axis < 0 || sc.rects[e.index].axis > 2) {
          error = "pt_scene: bad rect reference";
          return PT_ERR_INVALID_ARGUMENT;
        }
        const pt_rect& r = sc.rects[e.index];
        float axis_bits;
        const int32_t axis = r.axis;
        std::memcpy(&axis_bits, &axis, 4);
        planes[2 - r.axis].push_back(r.k);  // xy: k is z, xz: y, yz: x (rectangle.hpp:50,88,126)
        seg.rect.push_back(f4 { r.a0, r.a1, r.b0, r.b1 });
        seg.rect.push_back(f4 { r.k, axis_bits, 0.f, 0.f });
        seg.rect_aux.push_back(ObjAux { r.material, (int32_t)i });
        break;
      }
      case PT_HIT_TRIANGLE: {
        if (e.index < 0 || (uint32_t)e.index >= sc.n_triangles || !mat_ok(sc.triangles[e.index].material)) {
          error = "pt_scene: bad triangle reference";
          return PT_ERR_INVALID_ARGUMENT;
        }
        const pt_triangle& t = sc.triangles[e.index];
        const float e1[3] = { t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2] };
        const float e2[3] = { t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2] };
        seg.tri.push_back(f4 { t.v0[0], t.v0[1], t.v0[2], 0.f });
        seg.tri.push_back(f4 { e1[0], e1[1], e1[2], 0.f });
        seg.tri.push_back(f4 { e2[0], e2[1], e2[2], 0.f });
        TriAux a {};
        a.nx = e1[1] * e2[2] - e1[2] * e2[1];
        a.ny = e1[2] * e2[0] - e1[0] * e2[2];
        a.nz = e1[0] * e2[1] - e1[1] * e2[0];
        a.material = t.material, a.key = (int32_t)i;
        seg.tri_aux.push_back(a);
        break;
      }
      case PT_HIT_BOX: {
        if (e.index < 0 || (uint32_t)e.index >= sc.n_boxes || !mat_ok(sc.boxes[e.index].material)) {
          error = "pt_scene: bad box reference";
          return PT_ERR_INVALID_ARGUMENT;
        }
        const pt_box& bx = sc.boxes[e.index];
        for (int k = 0; k < 3; ++k) planes[k].push_back(bx.p0[k]), planes[k].push_back(bx.p1[k]);
        seg.box.push_back(f4 { bx.p0[0], bx.p0[1], bx.p0[2], 0.f });
        seg.box.push_back(f4 { bx.p1[0], bx.p1[1], bx.p1[2], 0.f });
        seg.box_aux.push_back(ObjAux { bx.material, (int32_t)i });
        break;
      }
      case PT_HIT_MEDIUM: {
        if (e.index < 0 || (uint32_t)e.index >= sc.n_media || !mat_ok(sc.media[e.index].material)) {
          error = "pt_scene: bad constant_medium reference";
          return PT_ERR_INVALID_ARGUMENT;
        }
        const pt_medium& m = sc.media[e.index];
        MediumRec r {};
        r.boundary_kind = m.boundary_kind;
        r.neg_inv_density = -1 / m.density;
        r.material = m.material, r.key = (int32_t)i;
        if (m.boundary_kind == PT_BOUNDARY_SPHERE) {
          if (m.boundary_index < 0 || (uint32_t)m.boundary_index >= sc.n_spheres) {
            error = "pt_scene: bad medium boundary sphere";
            return PT_ERR_INVALID_ARGUMENT;
          }
          const pt_sphere& s = sc.spheres[m.boundary_index];
          for (int k = 0; k < 3; ++k) r.c0[k] = s.center0[k], r.dv[k] = s.center1[k] - s.center0[k];
          r.radius = s.radius, r.r2 = s.radius * s.radius;
          r.time0 = s.time0, r.den = s.time1 - s.time0, r.moving = !(s.time0 == s.time1);
        } else if (m.boundary_kind == PT_BOUNDARY_BOX) {
          if (m.boundary_index < 0 || (uint32_t)m.boundary_index >= sc.n_boxes) {
            error = "pt_scene: bad medium boundary box";
            return PT_ERR_INVALID_ARGUMENT;
          }
          const pt_box& bx = sc.boxes[m.boundary_index];
          for (int k = 0; k < 3; ++k) r.p0[k] = bx.p0[k], r.p1[k] = bx.p1[k];
          for (int k = 0; k < 3; ++k) planes[k].push_back(bx.p0[k]), planes[k].push_back(bx.p1[k]);
        } else {
          error = "pt_scene: unknown medium boundary kind";
          return PT_ERR_INVALID_ARGUMENT;
        }
        b.flush(seg);
        Group g {};
        g.type = G_MEDIUM, g.begin = (int32_t)out.media.size(), g.count = 1, g.tree = g.gtree = -1;
        b.groups.push_back(g);
        out.media.push_back(r);
        break;
      }
      default:
        error = "pt_scene: unknown hittable kind";
        return PT_ERR_INVALID_ARGUMENT;
    }
  }
  b.flush(seg);

  if (out.sphere_aux.size() > kIdMask || out.moving_aux.size() > kIdMask || out.tri_aux.size() > kIdMask) {
    error = "pt_scene: too many objects";
    return PT_ERR_UNSUPPORTED;
  }

  out.n_groups = (uint32_t)b.groups.size();
  for (const Group& g : b.groups) {
    if ((g.type == G_SPHERE || g.type == G_MOVING_SPHERE) && out.n_media_groups != 0) ++out.n_late_sphere_groups;
    if (g.type == G_MEDIUM) ++out.n_media_groups;
    if (g.type == G_RECT || g.type == G_TRIANGLE || g.type == G_BOX) ++out.n_flat_groups;
  }
  out.n_objects = sc.n_hittables;
  out.off_groups = append(out.blob, b.groups.data(), b.groups.size() * sizeof(Group));
  out.off_sphere = append(out.blob, b.sph.data(), b.sph.size() * sizeof(f4));
  out.off_moving = append(out.blob, b.mov.data(), b.mov.size() * sizeof(f4));
  out.off_rect = append(out.blob, b.rect.data(), b.rect.size() * sizeof(f4));
  out.off_triangle = append(out.blob, b.tri.data(), b.tri.size() * sizeof(f4));
  out.off_box = append(out.blob, b.box.data(), b.box.size() * sizeof(f4));
  out.off_trees = append(out.blob, b.trees.data(), b.trees.size() * sizeof(Tree));
  out.off_nodes = append(out.blob, b.nodes.data(), b.nodes.size() * sizeof(f4));
  out.off_tree_ids = append(out.blob, b.tree_ids.data(), b.tree_ids.size() * sizeof(f4));
  out.n_trees = (uint32_t)b.trees.size();
  {
    std::vector<float> all;
    for (int k = 0; k < 3; ++k) {
      std::vector<float>& v = planes[k];
      v.erase(std::remove_if(v.begin(), v.end(), [](float x) { return x != x; }), v.end());  // (a NaN plane equals nothing)
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      out.n_planes[k] = (uint32_t)v.size();
      all.insert(all.end(), v.begin(), v.end());
    }
    out.off_planes = append(out.blob, all.data(), all.size() * sizeof(float));
  }
  out.flat_extent = (float)b.flat_extent;
  {
    // until compute_cull_boxes() has run for a camera, every chunk is scanned
    CullBoxes none;
    compute_cull_boxes(out, std::numeric_limits<float>::quiet_NaN(), std::numeric_limits<float>::quiet_NaN(), none);
    out.off_sphere_box = append(out.blob, none.sphere.data(), none.sphere.size() * sizeof(float));
    out.off_moving_box = append(out.blob, none.moving.data(), none.moving.size() * sizeof(float));
  }
  while (out.blob.size() % 16 || out.blob.empty()) out.blob.push_back(0);

  out.materials.assign(sc.materials, sc.materials + sc.n_materials);
  out.textures.assign(sc.textures, sc.textures + sc.n_textures);
  return PT_OK;
}

}  // namespace ptb
