// pt/flatten.hpp -- lower a std::vector<hittable_t> (pt/scene.hpp) to the flat, type-tagged
// pt_scene of include/pt_abi.h.  This is what replaces "wrap the variant vector in a sycl::buffer"
// (reference render.hpp:146-148): the vector order is kept in the ORDER table, every object gets
// its own material / texture rows, constant_medium boundaries go into the sphere / box arrays
// without an ORDER entry.
#ifndef PT_FLATTEN_HPP
#define PT_FLATTEN_HPP

#include <type_traits>
#include <variant>
#include <vector>

#include "pt/scene.hpp"
#include "pt_abi.h"
#include "ptscene_io.hpp"

namespace pt {

namespace detail {
inline void put3(float* dst, const sycl::float3& v) { dst[0] = v.x(), dst[1] = v.y(), dst[2] = v.z(); }

template <class... F> struct overloaded : F... { using F::operator()...; };
template <class... F> overloaded(F...) -> overloaded<F...>;
}  // namespace detail

class flattener {
 public:
  ptscene::owned_scene out;

  int texture(const texture_t& t) {
    pt_texture row {};
    row.kind = int(t.index());  // variant order == PT_TEX_*
    std::visit(detail::overloaded {
                   [&](const checker_texture& c) { detail::put3(row.color0, c.odd.rgb), detail::put3(row.color1, c.even.rgb); },
                   [&](const solid_texture& s) { detail::put3(row.color0, s.rgb); },
                   [&](const image_texture& i) {
                     row.width = std::uint32_t(i.width), row.height = std::uint32_t(i.height);
                     row.offset = i.offset, row.freq = i.cyclic_frequency;
                   } },
               t);
    out.textures.push_back(row);
    return int(out.textures.size()) - 1;
  }

  int material(const material_t& m) {
    pt_material row {};
    row.kind = int(m.index());  // variant order == PT_MAT_*
    row.texture = -1;
    std::visit(detail::overloaded {
                   [&](const lambertian_material& l) { row.texture = texture(l.albedo); },
                   [&](const metal_material& x) { detail::put3(row.albedo, x.albedo), row.param = x.fuzz; },
                   [&](const dielectric_material& d) { detail::put3(row.albedo, d.albedo), row.param = d.ref_idx; },
                   [&](const lightsource_material& l) { row.texture = texture(l.emit); },
                   [&](const isotropic_material& i) { row.texture = texture(i.albedo); } },
               m);
    out.materials.push_back(row);
    return int(out.materials.size()) - 1;
  }

  int add(const sphere& s) {
    pt_sphere row {};
    detail::put3(row.center0, s.center0), detail::put3(row.center1, s.center1);
    row.radius = s.radius, row.time0 = s.time0, row.time1 = s.time1, row.material = material(s.material_type);
    out.spheres.push_back(row);
    return int(out.spheres.size()) - 1;
  }
  template <int Axis> int add(const axis_rect<Axis>& r) {
    out.rects.push_back(pt_rect { r.a0, r.a1, r.b0, r.b1, r.k, Axis, material(r.material_type) });
    return int(out.rects.size()) - 1;
  }
  int add(const triangle& t) {
    pt_triangle row {};
    detail::put3(row.v0, t.v0), detail::put3(row.v1, t.v1), detail::put3(row.v2, t.v2);
    row.material = material(t.material_type);
    out.triangles.push_back(row);
    return int(out.triangles.size()) - 1;
  }
  int add(const box& b) {
    pt_box row {};
    detail::put3(row.p0, b.box_min), detail::put3(row.p1, b.box_max);
    row.material = material(b.material_type);
    out.boxes.push_back(row);
    return int(out.boxes.size()) - 1;
  }
  int add(const constant_medium& m) {
    pt_medium row {};
    row.boundary_kind = int(m.boundary.index());  // sphere = PT_BOUNDARY_SPHERE, box = PT_BOUNDARY_BOX
    row.boundary_index = std::visit([&](const auto& b) { return add(b); }, m.boundary);
    row.density = m.density;
    row.material = material(m.phase_function);
    out.media.push_back(row);
    return int(out.media.size()) - 1;
  }

  void top_level(const hittable_t& h) {
    pt_order_entry e {};
    e.kind = int(h.index());  // variant order == PT_HIT_*
    e.index = std::visit([&](const auto& obj) { return add(obj); }, h);
    out.order.push_back(e);
  }
};

// The whole vector, plus the image texel pool (image_texture::pool()).
inline ptscene::owned_scene flatten(const std::vector<hittable_t>& hittables) {
  flattener f;
  for (const auto& h : hittables) f.top_level(h);
  f.out.texture_bytes = image_texture::pool();
  return std::move(f.out);
}

inline pt_camera to_abi(const camera& c) {
  pt_camera o {};
  detail::put3(o.origin, c.origin), detail::put3(o.lower_left_corner, c.lower_left_corner);
  detail::put3(o.horizontal, c.horizontal), detail::put3(o.vertical, c.vertical);
  detail::put3(o.u, c.u), detail::put3(o.v, c.v), detail::put3(o.w, c.w);
  o.lens_radius = c.lens_radius, o.time0 = c.time0, o.time1 = c.time1;
  return o;
}

}  // namespace pt
#endif
