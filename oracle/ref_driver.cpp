// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libptref.so (see oracle/Makefile).
//
// The reference itself as an oracle: this translation unit includes the
// UNMODIFIED reference headers from /root/reference/include (in place, never
// copied) against the host shim in oracle/shim, and drives
//   * the reference's own entry point  render<W,H,S>()      (render.hpp:141-160)
//   * the reference's kernel body      render_pixel<W,H,S,D>() (render.hpp:25-106)
//     for arbitrary pixel subsets, seeded exactly as the executor does
//     (render.hpp:130-133).
// The scene arrives as the flat C-ABI pt_scene (include/pt_abi.h) and is
// rebuilt into a std::vector<hittable_t> with the reference's own
// constructors, so every value the reference derives at construction time
// (box sides, -1/density, clamped fuzz ...) is derived by reference code.
//
// width/height/samples/depth are template parameters in the reference, so only
// the instantiations listed in PTREF_CONFIGS exist here; the plain-C
// restatement (oracle/pt_oracle.c) covers arbitrary sizes and is proven
// bit-identical against this library by tests/test_oracle.py.
#include <sycl.hpp>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <typeinfo>

#include <omp.h>

#include "render.hpp"  // the reference, in place

#include "pt_abi.h"

static_assert(sizeof(sycl::float3) == 12, "float3 must be 12 bytes");
static_assert(sizeof(camera) == sizeof(pt_camera) && sizeof(pt_camera) == 96,
              "pt_camera mirrors camera.hpp:21-46");
static_assert(std::is_trivially_copyable_v<camera>);

namespace {

thread_local std::string g_error;

// ---- image provider for the stb shim ------------------------------------
struct pending_image {
  const std::uint8_t* data;
  int w, h;
};
std::map<std::string, pending_image>& pending_images() {
  static std::map<std::string, pending_image> m;
  return m;
}
const std::uint8_t* provide_image(const char* name, int* w, int* h) {
  auto it = pending_images().find(name);
  if (it == pending_images().end()) return nullptr;
  *w = it->second.w;
  *h = it->second.h;
  return it->second.data;
}

color col(const float* p) { return color { p[0], p[1], p[2] }; }
point pnt(const float* p) { return point { p[0], p[1], p[2] }; }

struct rebuilt_scene {
  std::vector<hittable_t> hittables;
};

// The reference's texture pool is a private static that only grows
// (texture.hpp:71,113-114); identical images are registered once per process.
std::map<std::string, texture_t>& image_cache() {
  static std::map<std::string, texture_t> m;
  return m;
}

bool make_texture(const pt_scene& s, int idx, texture_t& out) {
  if (idx < 0 || (std::uint32_t)idx >= s.n_textures) {
    g_error = "texture index out of range";
    return false;
  }
  const pt_texture& t = s.textures[idx];
  switch (t.kind) {
    case PT_TEX_CHECKER:
      out = checker_texture(col(t.color0), col(t.color1));
      return true;
    case PT_TEX_SOLID:
      out = solid_texture(col(t.color0));
      return true;
    case PT_TEX_IMAGE: {
      const std::size_t nbytes = (std::size_t)t.width * t.height * 3u;
      const bool fallback = (t.offset == 0);  // texture.hpp:105-111: failed load -> 1x1 at texel 0
      if (!fallback && (t.offset * 3u + nbytes > s.n_texture_bytes)) {
        g_error = "image texture exceeds the texture byte pool";
        return false;
      }
      // key: content + dims + frequency
      std::string key;
      if (!fallback) key.assign((const char*)s.texture_bytes + t.offset * 3u, nbytes);
      key += "|" + std::to_string(t.width) + "x" + std::to_string(t.height) + "@" +
             std::to_string(t.freq) + (fallback ? "F" : "");
      auto it = image_cache().find(key);
      if (it == image_cache().end()) {
        ptref_shim::image_provider() = provide_image;
        const std::string name = "mem:" + std::to_string(image_cache().size());
        if (!fallback)
          pending_images()[name] = { s.texture_bytes + t.offset * 3u, (int)t.width, (int)t.height };
        texture_t made = image_texture::image_texture_factory(name.c_str(), t.freq);
        pending_images().erase(name);
        it = image_cache().emplace(key, made).first;
      }
      out = it->second;
      return true;
    }
  }
  g_error = "unknown texture kind";
  return false;
}

bool make_material(const pt_scene& s, int idx, material_t& out) {
  if (idx < 0 || (std::uint32_t)idx >= s.n_materials) {
    g_error = "material index out of range";
    return false;
  }
  const pt_material& m = s.materials[idx];
  texture_t tex;
  switch (m.kind) {
    case PT_MAT_LAMBERTIAN:
      if (!make_texture(s, m.texture, tex)) return false;
      out = lambertian_material(tex);
      return true;
    case PT_MAT_METAL:
      out = metal_material(col(m.albedo), m.param);
      return true;
    case PT_MAT_DIELECTRIC:
      out = dielectric_material(m.param, col(m.albedo));
      return true;
    case PT_MAT_LIGHTSOURCE:
      if (!make_texture(s, m.texture, tex)) return false;
      out = lightsource_material(tex);
      return true;
    case PT_MAT_ISOTROPIC:
      if (!make_texture(s, m.texture, tex)) return false;
      out = isotropic_material(tex);
      return true;
  }
  g_error = "unknown material kind";
  return false;
}

bool make_sphere(const pt_scene& s, int idx, sphere& out) {
  if (idx < 0 || (std::uint32_t)idx >= s.n_spheres) {
    g_error = "sphere index out of range";
    return false;
  }
  const pt_sphere& p = s.spheres[idx];
  material_t m;
  if (!make_material(s, p.material, m)) return false;
  out = sphere(pnt(p.center0), pnt(p.center1), p.time0, p.time1, p.radius, m);
  return true;
}

bool make_box(const pt_scene& s, int idx, box& out) {
  if (idx < 0 || (std::uint32_t)idx >= s.n_boxes) {
    g_error = "box index out of range";
    return false;
  }
  const pt_box& b = s.boxes[idx];
  material_t m;
  if (!make_material(s, b.material, m)) return false;
  out = box(pnt(b.p0), pnt(b.p1), m);
  return true;
}

bool rebuild(const pt_scene& s, rebuilt_scene& out) {
  out.hittables.clear();
  out.hittables.reserve(s.n_hittables);
  for (std::uint32_t i = 0; i < s.n_hittables; ++i) {
    const pt_order_entry& e = s.order[i];
    material_t m;
    switch (e.kind) {
      case PT_HIT_SPHERE: {
        sphere sp;
        if (!make_sphere(s, e.index, sp)) return false;
        out.hittables.emplace_back(sp);
        break;
      }
      case PT_HIT_RECT: {
        if (e.index < 0 || (std::uint32_t)e.index >= s.n_rects) {
          g_error = "rect index out of range";
          return false;
        }
        const pt_rect& r = s.rects[e.index];
        if (r.axis != PT_AXIS_XY) {
          g_error = "the reference's hittable_t only holds xy_rect at top level (render.hpp:22-23)";
          return false;
        }
        if (!make_material(s, r.material, m)) return false;
        out.hittables.emplace_back(xy_rect(r.a0, r.a1, r.b0, r.b1, r.k, m));
        break;
      }
      case PT_HIT_TRIANGLE: {
        if (e.index < 0 || (std::uint32_t)e.index >= s.n_triangles) {
          g_error = "triangle index out of range";
          return false;
        }
        const pt_triangle& t = s.triangles[e.index];
        if (!make_material(s, t.material, m)) return false;
        out.hittables.emplace_back(triangle(pnt(t.v0), pnt(t.v1), pnt(t.v2), m));
        break;
      }
      case PT_HIT_BOX: {
        box b;
        if (!make_box(s, e.index, b)) return false;
        out.hittables.emplace_back(b);
        break;
      }
      case PT_HIT_MEDIUM: {
        if (e.index < 0 || (std::uint32_t)e.index >= s.n_media) {
          g_error = "medium index out of range";
          return false;
        }
        const pt_medium& md = s.media[e.index];
        if (md.material < 0 || (std::uint32_t)md.material >= s.n_materials ||
            s.materials[md.material].kind != PT_MAT_ISOTROPIC) {
          g_error = "constant_medium needs an isotropic phase function";
          return false;
        }
        texture_t tex;
        if (!make_texture(s, s.materials[md.material].texture, tex)) return false;
        if (md.boundary_kind == PT_BOUNDARY_SPHERE) {
          sphere sp;
          if (!make_sphere(s, md.boundary_index, sp)) return false;
          out.hittables.emplace_back(constant_medium { sp, md.density, tex });
        } else {
          box b;
          if (!make_box(s, md.boundary_index, b)) return false;
          out.hittables.emplace_back(constant_medium { b, md.density, tex });
        }
        break;
      }
      default:
        g_error = "unknown hittable kind";
        return false;
    }
  }
  return true;
}

camera camera_from_abi(const pt_camera& c) {
  camera cam { point { 0, 0, 1 }, point { 0, 0, 0 }, vec { 0, 1, 0 }, 40.f, 1.f, 0.f, 1.f };
  std::memcpy(static_cast<void*>(&cam), &c, sizeof(pt_camera));
  return cam;
}

// Framebuffer adapter with the access pattern render_pixel uses:
// fb_acc[y][x] = colour (render.hpp:105), mapped onto a pt_region.
struct region_fb {
  float* out;
  std::int64_t pitch;
  pt_region rg;
  struct row {
    float* base;
    int x0;
    color& operator[](std::size_t x) const {
      return *reinterpret_cast<color*>(base + 3 * ((long)x - x0));
    }
  };
  row operator[](std::size_t y) const {
    const long k = ((long)y - rg.y0) / rg.y_stride;
    return row { out + k * pitch, rg.x0 };
  }
};

template <int W, int H, int S, int D>
void run_region(const camera& cam, std::vector<hittable_t>& hittables, const pt_region& rg,
                float* out, std::int64_t pitch) {
  sycl::buffer<hittable_t, 1> hbuf(hittables.data(), sycl::range<1>(hittables.size()));
  auto tbuf = image_texture::freeze();
  sycl::handler cgh;
  auto hacc = hbuf.get_access<sycl::access::mode::read>(cgh);
  auto tacc = tbuf.get_access<sycl::access::mode::read>(cgh);
  region_fb fb { out, pitch, rg };
  const long npix = (long)rg.w * rg.h;
#pragma omp parallel for schedule(runtime)
  for (long i = 0; i < npix; ++i) {
    const int x = rg.x0 + (int)(i % rg.w);
    const int y = rg.y0 + (int)(i / rg.w) * rg.y_stride;
    // render.hpp:130-133
    auto init_generator_state = std::hash<std::size_t> {}((std::size_t)y * W + (std::size_t)x);
    LocalPseudoRNG rng(init_generator_state);
    task_context ctx { rng, tacc.get_pointer() };
    render_pixel<W, H, S, D>(ctx, x, y, cam, hacc, fb);
  }
}

template <int W, int H, int S>
void run_full(camera& cam, std::vector<hittable_t>& hittables, float* fb_out) {
  sycl::queue q;
  sycl::buffer<color, 2> fb(sycl::range<2>(H, W));
  render<W, H, S>(q, fb, hittables, cam);  // the reference entry point, depth = 50
  std::memcpy(fb_out, fb.host_data(), sizeof(float) * 3u * W * H);
}

// (width, height, samples, depth) instantiations of the reference kernel.
#define PTREF_CONFIGS(X)                                                           \
  X(800, 480, 100, 50) X(800, 480, 32, 50) X(800, 480, 8, 50) X(800, 480, 1, 50)  \
  X(1920, 1080, 64, 50) X(1920, 1080, 256, 50) X(1024, 1024, 1024, 50)            \
  X(3840, 2160, 4096, 50)                                                          \
  X(200, 120, 16, 50) X(200, 120, 4, 50) X(200, 120, 64, 50)                       \
  X(64, 48, 1, 50) X(64, 48, 4, 50) X(64, 48, 32, 50) X(64, 48, 16, 3) X(64, 48, 16, 1) \
  X(96, 64, 256, 50) X(33, 17, 5, 7)

void apply_schedule(int dynamic, int nthreads) {
  omp_set_schedule(dynamic ? omp_sched_dynamic : omp_sched_static, dynamic ? 1 : 0);
  if (nthreads > 0) omp_set_num_threads(nthreads);
}

}  // namespace

extern "C" {

const char* ptref_last_error() { return g_error.c_str(); }

int ptref_supported(int w, int h, int spp, int depth) {
#define X(W, H, S, D) \
  if (w == W && h == H && spp == S && depth == D) return 1;
  PTREF_CONFIGS(X)
#undef X
  return 0;
}

// Writes up to `cap` quadruples (w,h,spp,depth); returns the number of configs.
int ptref_list_configs(int* out, int cap) {
  int n = 0;
#define X(W, H, S, D)      \
  if (n < cap) {           \
    out[4 * n + 0] = W;    \
    out[4 * n + 1] = H;    \
    out[4 * n + 2] = S;    \
    out[4 * n + 3] = D;    \
  }                        \
  ++n;
  PTREF_CONFIGS(X)
#undef X
  return n;
}

int ptref_max_threads() { return omp_get_max_threads(); }

// render_pixel<> over a region with the executor's seeding.
int ptref_render_region(int w, int h, int spp, int depth, const pt_camera* cam,
                        const pt_scene* scene, const pt_region* region, float* out,
                        std::int64_t out_row_pitch, int dynamic_schedule, int nthreads) {
  if (!cam || !scene || !region || !out) {
    g_error = "null argument";
    return -1;
  }
  rebuilt_scene rs;
  if (!rebuild(*scene, rs)) return -1;
  camera c = camera_from_abi(*cam);
  apply_schedule(dynamic_schedule, nthreads);
#define X(W, H, S, D)                                                        \
  if (w == W && h == H && spp == S && depth == D) {                          \
    run_region<W, H, S, D>(c, rs.hittables, *region, out, out_row_pitch);    \
    return 0;                                                                \
  }
  PTREF_CONFIGS(X)
#undef X
  g_error = "configuration not instantiated in libptref (width/height/samples/depth are "
            "template parameters of the reference kernel)";
  return -4;
}

// The reference's own render<W,H,S>() (depth is its internal constexpr 50).
int ptref_render_full(int w, int h, int spp, const pt_camera* cam, const pt_scene* scene,
                      float* fb, int dynamic_schedule, int nthreads) {
  if (!cam || !scene || !fb) {
    g_error = "null argument";
    return -1;
  }
  rebuilt_scene rs;
  if (!rebuild(*scene, rs)) return -1;
  camera c = camera_from_abi(*cam);
  apply_schedule(dynamic_schedule, nthreads);
#define X(W, H, S, D)                          \
  if (w == W && h == H && spp == S && D == 50) { \
    run_full<W, H, S>(c, rs.hittables, fb);    \
    return 0;                                  \
  }
  PTREF_CONFIGS(X)
#undef X
  g_error = "configuration not instantiated in libptref";
  return -4;
}

// Camera constructor arithmetic (camera.hpp:67-87), run by the reference.
void ptref_make_camera(const float look_from[3], const float look_at[3], const float vup[3],
                       float vfov_deg, float aspect, float aperture, float focus_dist, float t0,
                       float t1, pt_camera* out) {
  camera cam { pnt(look_from), pnt(look_at), pnt(vup), vfov_deg, aspect, aperture, focus_dist,
               t0,             t1 };
  std::memcpy(out, static_cast<const void*>(&cam), sizeof(pt_camera));
}

// ---- known-answer generators (xorshift.hpp:64-93, rtweekend.hpp:33-92) ----
void ptref_kat_xorshift(std::uint32_t seed, int n, std::uint32_t* out) {
  xorshift<> g { seed };
  for (int i = 0; i < n; ++i) out[i] = g();
}
void ptref_kat_float(std::uint32_t seed, int n, float* out) {
  LocalPseudoRNG r { seed };
  for (int i = 0; i < n; ++i) out[i] = r.float_t();
}
// kind: 0 unit_vec, 1 in_unit_ball, 2 in_unit_disk, 3 vec_t; n vectors -> 3n floats
void ptref_kat_vec(std::uint32_t seed, int kind, int n, float* out) {
  LocalPseudoRNG r { seed };
  for (int i = 0; i < n; ++i) {
    vec v = kind == 0 ? r.unit_vec() : kind == 1 ? r.in_unit_ball() : kind == 2 ? r.in_unit_disk() : r.vec_t();
    out[3 * i] = v.x();
    out[3 * i + 1] = v.y();
    out[3 * i + 2] = v.z();
  }
}
// camera::get_ray (camera.hpp:93-100): n rays from one stream -> 7n floats (o, d, time)
void ptref_kat_get_ray(const pt_camera* cam, std::uint32_t seed, int n, const float* st, float* out) {
  camera c = camera_from_abi(*cam);
  LocalPseudoRNG r { seed };
  for (int i = 0; i < n; ++i) {
    ray ry = c.get_ray(st[2 * i], st[2 * i + 1], r);
    out[7 * i + 0] = ry.origin().x();
    out[7 * i + 1] = ry.origin().y();
    out[7 * i + 2] = ry.origin().z();
    out[7 * i + 3] = ry.direction().x();
    out[7 * i + 4] = ry.direction().y();
    out[7 * i + 5] = ry.direction().z();
    out[7 * i + 6] = ry.time();
  }
}

// Closest-hit of ONE ray against the scene with the reference's sequential
// scan (render.hpp:30-51) + the scatter of the hit material.  out[0..]:
// hit(0/1), t, p(3), normal(3), front_face, u, v, scattered(0/1), att(3),
// scat_o(3), scat_d(3), emitted(3), rng_state_after  = 25 floats/u32.
int ptref_kat_hit_scatter(const pt_scene* scene, const float* ray7, std::uint32_t seed, float* out) {
  rebuilt_scene rs;
  if (!rebuild(*scene, rs)) return -1;
  auto tbuf = image_texture::freeze();
  sycl::handler cgh;
  auto tacc = tbuf.get_access<sycl::access::mode::read>(cgh);
  LocalPseudoRNG rng(seed);
  task_context ctx { rng, tacc.get_pointer() };
  ray r { point { ray7[0], ray7[1], ray7[2] }, vec { ray7[3], ray7[4], ray7[5] }, ray7[6] };
  hit_record rec {}, temp_rec {};
  material_t mat, temp_mat;
  bool hit_anything = false;
  float closest = infinity;
  for (auto& h : rs.hittables) {
    if (dev_visit([&](auto&& arg) { return arg.hit(ctx, r, 0.001f, closest, temp_rec, temp_mat); },
                  h)) {
      hit_anything = true;
      closest = temp_rec.t;
      rec = temp_rec;
      mat = temp_mat;
    }
  }
  std::memset(out, 0, 26 * sizeof(float));
  out[0] = hit_anything;
  if (hit_anything) {
    out[1] = rec.t;
    out[2] = rec.p.x(), out[3] = rec.p.y(), out[4] = rec.p.z();
    out[5] = rec.normal.x(), out[6] = rec.normal.y(), out[7] = rec.normal.z();
    out[8] = rec.front_face;
    out[9] = rec.u, out[10] = rec.v;
    color att { 1.f, 1.f, 1.f };
    ray scattered;
    color emitted = dev_visit([&](auto&& arg) { return arg.emitted(ctx, rec); }, mat);
    bool sc = dev_visit([&](auto&& arg) { return arg.scatter(ctx, r, rec, att, scattered); }, mat);
    out[11] = sc;
    out[12] = att.x(), out[13] = att.y(), out[14] = att.z();
    if (sc) {
      out[15] = scattered.origin().x(), out[16] = scattered.origin().y(), out[17] = scattered.origin().z();
      out[18] = scattered.direction().x(), out[19] = scattered.direction().y(),
      out[20] = scattered.direction().z();
    }
    out[21] = emitted.x(), out[22] = emitted.y(), out[23] = emitted.z();
    out[24] = (float)mat.index();
  }
  return 0;
}

}  // extern "C"
