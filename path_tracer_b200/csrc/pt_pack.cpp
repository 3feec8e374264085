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

struct MovingClass {
  float time0, time1;
  std::vector<f4> data;  // 2 per sphere
  std::vector<SphereAux> aux;
};

struct Segment {
  std::vector<f4> sph;
  std::vector<SphereAux> sph_aux;
  std::vector<MovingClass> classes;
  std::vector<f4> rect;
  std::vector<ObjAux> rect_aux;
  std::vector<f4> tri;
  std::vector<TriAux> tri_aux;
  std::vector<f4> box;
  std::vector<ObjAux> box_aux;
  bool empty() const { return sph.empty() && classes.empty() && rect.empty() && tri.empty() && box.empty(); }
};

struct Builder {
  std::vector<Group> groups;
  std::vector<f4> sph, mov, rect, tri, box;
  PackedScene& out;
  explicit Builder(PackedScene& o) : out(o) {}

  static void pad_spheres(std::vector<f4>& data, std::vector<SphereAux>& aux, int per_sphere) {
    // A padding sphere can never become a candidate: r*r = -inf makes
    // c = dot(oc,oc) - r*r = +inf, so discriminant = b*b - a*c is -inf or NaN.
    const float ninf = -std::numeric_limits<float>::infinity();
    while (aux.size() % kSphereChunk) {
      data.push_back(f4 { 0.f, 0.f, 0.f, ninf });
      if (per_sphere == 2) data.push_back(f4 { 0.f, 0.f, 0.f, 0.f });
      SphereAux a {};
      a.radius = 1.f, a.material = -1, a.key = std::numeric_limits<int32_t>::min();
      aux.push_back(a);
    }
  }

  void flush(Segment& s) {
    if (!s.sph.empty()) {
      pad_spheres(s.sph, s.sph_aux, 1);
      Group g {};
      g.type = G_SPHERE, g.begin = (int32_t)out.sphere_aux.size(), g.count = (int32_t)s.sph_aux.size();
      groups.push_back(g);
      sph.insert(sph.end(), s.sph.begin(), s.sph.end());
      out.sphere_aux.insert(out.sphere_aux.end(), s.sph_aux.begin(), s.sph_aux.end());
    }
    for (auto& c : s.classes) {
      pad_spheres(c.data, c.aux, 2);
      Group g {};
      g.type = G_MOVING_SPHERE, g.begin = (int32_t)out.moving_aux.size(), g.count = (int32_t)c.aux.size();
      g.time0 = c.time0, g.den = c.time1 - c.time0;
      groups.push_back(g);
      mov.insert(mov.end(), c.data.begin(), c.data.end());
      out.moving_aux.insert(out.moving_aux.end(), c.aux.begin(), c.aux.end());
    }
    if (!s.rect_aux.empty()) {
      Group g {};
      g.type = G_RECT, g.begin = (int32_t)out.rect_aux.size(), g.count = (int32_t)s.rect_aux.size();
      groups.push_back(g);
      rect.insert(rect.end(), s.rect.begin(), s.rect.end());
      out.rect_aux.insert(out.rect_aux.end(), s.rect_aux.begin(), s.rect_aux.end());
    }
    if (!s.tri_aux.empty()) {
      Group g {};
      g.type = G_TRIANGLE, g.begin = (int32_t)out.tri_aux.size(), g.count = (int32_t)s.tri_aux.size();
      groups.push_back(g);
      tri.insert(tri.end(), s.tri.begin(), s.tri.end());
      out.tri_aux.insert(out.tri_aux.end(), s.tri_aux.begin(), s.tri_aux.end());
    }
    if (!s.box_aux.empty()) {
      Group g {};
      g.type = G_BOX, g.begin = (int32_t)out.box_aux.size(), g.count = (int32_t)s.box_aux.size();
      groups.push_back(g);
      box.insert(box.end(), s.box.begin(), s.box.end());
      out.box_aux.insert(out.box_aux.end(), s.box_aux.begin(), s.box_aux.end());
    }
    s = Segment {};
  }
};

uint32_t append(std::vector<unsigned char>& blob, const void* p, size_t bytes) {
  while (blob.size() % 16) blob.push_back(0);
  const uint32_t off = (uint32_t)blob.size();
  const unsigned char* b = static_cast<const unsigned char*>(p);
  blob.insert(blob.end(), b, b + bytes);
  return off;
}

}  // namespace

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

  Builder b(out);
  Segment seg;
  for (uint32_t i = 0; i < sc.n_hittables; ++i) {
    const pt_order_entry e = sc.order[i];
    switch (e.kind) {
      case PT_HIT_SPHERE: {
        if (e.index < 0 || (uint32_t)e.index >= sc.n_spheres || !mat_ok(sc.spheres[e.index].material)) {
          error = "pt_scene: bad sphere reference";
          return PT_ERR_INVALID_ARGUMENT;
        }
        const pt_sphere& s = sc.spheres[e.index];
        SphereAux a {};
        a.radius = s.radius, a.material = s.material, a.key = -1 - (int32_t)i;
        // The scan blob carries r*r inflated by the miss filter's margin, rounded up (pt_kernel.cu,
        // "conservative miss filter"); the exact r*r is recomputed from the side table's radius.
        const float r2_exact = s.radius * s.radius;
        const float r2 = std::nextafter(r2_exact * (1.0f + 2.0f * 4.0e-6f / (1.0f - 4.0e-6f)),
                                        std::numeric_limits<float>::infinity());
        if (s.time0 == s.time1) {  // sphere.hpp:52
          seg.sph.push_back(f4 { s.center0[0], s.center0[1], s.center0[2], r2 });
          seg.sph_aux.push_back(a);
        } else {
          MovingClass* cls = nullptr;
          for (auto& c : seg.classes)
            if (c.time0 == s.time0 && c.time1 == s.time1) cls = &c;
          if (!cls) {
            seg.classes.push_back(MovingClass { s.time0, s.time1, {}, {} });
            cls = &seg.classes.back();
          }
          a.time0 = s.time0, a.den = s.time1 - s.time0;
          cls->data.push_back(f4 { s.center0[0], s.center0[1], s.center0[2], r2 });
          cls->data.push_back(f4 { s.center1[0] - s.center0[0], s.center1[1] - s.center0[1],
                                   s.center1[2] - s.center0[2], 0.f });
          cls->aux.push_back(a);
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
        } else {
          error = "pt_scene: unknown medium boundary kind";
          return PT_ERR_INVALID_ARGUMENT;
        }
        b.flush(seg);
        Group g {};
        g.type = G_MEDIUM, g.begin = (int32_t)out.media.size(), g.count = 1;
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
  out.n_objects = sc.n_hittables;
  out.off_groups = append(out.blob, b.groups.data(), b.groups.size() * sizeof(Group));
  out.off_sphere = append(out.blob, b.sph.data(), b.sph.size() * sizeof(f4));
  out.off_moving = append(out.blob, b.mov.data(), b.mov.size() * sizeof(f4));
  out.off_rect = append(out.blob, b.rect.data(), b.rect.size() * sizeof(f4));
  out.off_triangle = append(out.blob, b.tri.data(), b.tri.size() * sizeof(f4));
  out.off_box = append(out.blob, b.box.data(), b.box.size() * sizeof(f4));
  while (out.blob.size() % 16 || out.blob.empty()) out.blob.push_back(0);

  out.materials.assign(sc.materials, sc.materials + sc.n_materials);
  out.textures.assign(sc.textures, sc.textures + sc.n_textures);
  return PT_OK;
}

}  // namespace ptb
