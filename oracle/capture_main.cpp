// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/capture_main (see oracle/Makefile).
//
// Runs the UNMODIFIED reference application /root/reference/src/main.cpp
// (included in place, `main` renamed by the preprocessor) on top of the oracle
// shim, and observes what it hands to render<>():
//   * the std::vector<hittable_t> it built (main.cpp:67-161), seen when
//     render() wraps it in a sycl::buffer (render.hpp:146-147);
//   * the image-texture byte pool (texture.hpp:126-131);
//   * the camera: main.cpp:168-183 passes literals to the reference
//     constructor; the same literals are passed here and the result is checked
//     to be byte-identical to the camera captured inside the kernel closure.
// The flattened scene is written as a PTSCENE1 file (tests/golden/c1_scene.ptsc
// is produced this way by oracle/gen_golden.py).  With --run the kernel is
// executed as well and the 8-bit image main.cpp would have written to out.png
// (main.cpp:33-59) is saved raw, for end-to-end drop-in comparison.
#include <sycl.hpp>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define main ptref_reference_main
#include "main.cpp"  // -I /root/reference/src
#undef main

#include "pt_abi.h"
#include "ptscene_io.hpp"

namespace {

struct image_texture_mirror {  // texture.hpp:75-81
  std::size_t width, height, offset;
  float cyclic_frequency;
};
static_assert(sizeof(image_texture_mirror) == sizeof(image_texture));

const hittable_t* g_hittables = nullptr;
std::size_t g_n_hittables = 0;
const std::uint8_t* g_pool = nullptr;
std::size_t g_pool_bytes = 0;
std::string g_out_path;
std::string g_png_path;
bool g_run = false;
int g_status = 1;

void on_buffer(const std::type_info& t, const void* p, std::size_t n) {
  if (t == typeid(hittable_t)) {
    g_hittables = static_cast<const hittable_t*>(p);
    g_n_hittables = n;
  } else if (t == typeid(std::uint8_t)) {
    g_pool = static_cast<const std::uint8_t*>(p);
    g_pool_bytes = n;
  }
}

struct flat_builder {
  ptscene::owned_scene s;

  int add_texture(const texture_t& t) {
    pt_texture o {};
    int dummy_ctx = 0;
    hit_record rec {};
    if (auto* c = std::get_if<checker_texture>(&t)) {
      o.kind = PT_TEX_CHECKER;
      color odd = c->odd.value(dummy_ctx, rec), even = c->even.value(dummy_ctx, rec);
      o.color0[0] = odd.x(), o.color0[1] = odd.y(), o.color0[2] = odd.z();
      o.color1[0] = even.x(), o.color1[1] = even.y(), o.color1[2] = even.z();
    } else if (auto* so = std::get_if<solid_texture>(&t)) {
      o.kind = PT_TEX_SOLID;
      color v = so->value(dummy_ctx, rec);
      o.color0[0] = v.x(), o.color0[1] = v.y(), o.color0[2] = v.z();
    } else {
      const image_texture& im = std::get<image_texture>(t);
      image_texture_mirror m;
      std::memcpy(&m, static_cast<const void*>(&im), sizeof m);
      o.kind = PT_TEX_IMAGE;
      o.width = (std::uint32_t)m.width, o.height = (std::uint32_t)m.height;
      o.offset = m.offset, o.freq = m.cyclic_frequency;
    }
    s.textures.push_back(o);
    return (int)s.textures.size() - 1;
  }

  int add_material(const material_t& m) {
    pt_material o {};
    o.kind = (int)m.index();
    o.texture = -1;
    if (auto* l = std::get_if<lambertian_material>(&m))
      o.texture = add_texture(l->albedo);
    else if (auto* me = std::get_if<metal_material>(&m)) {
      o.albedo[0] = me->albedo.x(), o.albedo[1] = me->albedo.y(), o.albedo[2] = me->albedo.z();
      o.param = me->fuzz;
    } else if (auto* d = std::get_if<dielectric_material>(&m)) {
      o.albedo[0] = d->albedo.x(), o.albedo[1] = d->albedo.y(), o.albedo[2] = d->albedo.z();
      o.param = d->ref_idx;
    } else if (auto* li = std::get_if<lightsource_material>(&m))
      o.texture = add_texture(li->emit);
    else
      o.texture = add_texture(std::get<isotropic_material>(m).albedo);
    s.materials.push_back(o);
    return (int)s.materials.size() - 1;
  }

  int add_sphere(const sphere& sp) {
    pt_sphere o {};
    o.center0[0] = sp.center0.x(), o.center0[1] = sp.center0.y(), o.center0[2] = sp.center0.z();
    o.center1[0] = sp.center1.x(), o.center1[1] = sp.center1.y(), o.center1[2] = sp.center1.z();
    o.radius = sp.radius, o.time0 = sp.time0, o.time1 = sp.time1;
    o.material = add_material(sp.material_type);
    s.spheres.push_back(o);
    return (int)s.spheres.size() - 1;
  }

  int add_box(const box& b) {
    pt_box o {};
    o.p0[0] = b.box_min.x(), o.p0[1] = b.box_min.y(), o.p0[2] = b.box_min.z();
    o.p1[0] = b.box_max.x(), o.p1[1] = b.box_max.y(), o.p1[2] = b.box_max.z();
    o.material = add_material(b.material_type);
    s.boxes.push_back(o);
    return (int)s.boxes.size() - 1;
  }

  void add(const hittable_t& h) {
    pt_order_entry e {};
    e.kind = (int)h.index();
    if (auto* sp = std::get_if<sphere>(&h))
      e.index = add_sphere(*sp);
    else if (auto* r = std::get_if<xy_rect>(&h)) {
      pt_rect o { r->x0, r->x1, r->y0, r->y1, r->k, PT_AXIS_XY, add_material(r->material_type) };
      s.rects.push_back(o);
      e.index = (int)s.rects.size() - 1;
    } else if (auto* t = std::get_if<triangle>(&h)) {
      pt_triangle o {};
      o.v0[0] = t->v0.x(), o.v0[1] = t->v0.y(), o.v0[2] = t->v0.z();
      o.v1[0] = t->v1.x(), o.v1[1] = t->v1.y(), o.v1[2] = t->v1.z();
      o.v2[0] = t->v2.x(), o.v2[1] = t->v2.y(), o.v2[2] = t->v2.z();
      o.material = add_material(t->material_type);
      s.triangles.push_back(o);
      e.index = (int)s.triangles.size() - 1;
    } else if (auto* b = std::get_if<box>(&h))
      e.index = add_box(*b);
    else {
      const constant_medium& cm = std::get<constant_medium>(h);
      pt_medium o {};
      if (auto* bs = std::get_if<sphere>(&cm.boundary)) {
        o.boundary_kind = PT_BOUNDARY_SPHERE;
        o.boundary_index = add_sphere(*bs);
      } else {
        o.boundary_kind = PT_BOUNDARY_BOX;
        o.boundary_index = add_box(std::get<box>(cm.boundary));
      }
      // constant_medium.hpp:20 stores -1/d; d itself is not kept.  -1/(-1/d) is
      // exact for the powers of two and checked below for everything else.
      o.density = -1 / cm.neg_inv_density;
      if (-1 / o.density != cm.neg_inv_density) {
        std::fprintf(stderr, "capture: density does not round-trip\n");
        std::exit(3);
      }
      o.material = add_material(cm.phase_function);
      s.media.push_back(o);
      e.index = (int)s.media.size() - 1;
    }
    s.order.push_back(e);
  }
};

bool on_kernel(const void* closure, std::size_t closure_bytes, std::size_t rows, std::size_t cols) {
  if (!g_hittables || !g_pool) {
    std::fprintf(stderr, "capture: buffers not observed\n");
    std::exit(2);
  }
  flat_builder fb;
  for (std::size_t i = 0; i < g_n_hittables; ++i) fb.add(g_hittables[i]);
  fb.s.texture_bytes.assign(g_pool, g_pool + g_pool_bytes);

  // main.cpp:168-183, literal for literal, through the reference constructor.
  point look_from { 13, 3, 3 };
  point look_at { 0, -1, 0 };
  vec vup { 0, 1, 0 };
  real_t angle = 40;
  real_t aperture = 0.04f;
  real_t focus_dist = length(look_at - look_from);
  camera cam { look_from, look_at,    vup,  angle, static_cast<real_t>(cols) / rows,
               aperture,  focus_dist, 0.0f, 1.0f };
  // ...and it must be the very camera the kernel closure captured by value.
  const char* cb = static_cast<const char*>(closure);
  bool found = false;
  for (std::size_t off = 0; off + sizeof(camera) <= closure_bytes; ++off)
    if (std::memcmp(cb + off, static_cast<const void*>(&cam), sizeof(camera)) == 0) found = true;
  if (!found) {
    std::fprintf(stderr, "capture: restated camera literals do not match the kernel's camera\n");
    std::exit(4);
  }
  pt_camera pc;
  std::memcpy(&pc, static_cast<const void*>(&cam), sizeof pc);
  ptscene::file_meta meta { (int)cols, (int)rows, 100 /* main.cpp:186 */, 50 /* render.hpp:144 */ };
  if (!ptscene::save(g_out_path.c_str(), fb.s, pc, meta)) {
    std::fprintf(stderr, "capture: cannot write %s\n", g_out_path.c_str());
    std::exit(5);
  }
  std::fprintf(stderr, "capture: %zu hittables (%zu spheres, %zu rects, %zu triangles, %zu boxes, %zu media), "
                       "%zu texture bytes -> %s\n",
               fb.s.order.size(), fb.s.spheres.size(), fb.s.rects.size(), fb.s.triangles.size(),
               fb.s.boxes.size(), fb.s.media.size(), fb.s.texture_bytes.size(), g_out_path.c_str());
  g_status = 0;
  return !g_run;  // true = skip the (minute-long) render
}

std::string g_image_dir;
std::vector<std::vector<std::uint8_t>> g_images;
const std::uint8_t* provide(const char* name, int* w, int* h) {
  std::string base = name;
  auto slash = base.find_last_of('/');
  if (slash != std::string::npos) base = base.substr(slash + 1);
  const std::string path = g_image_dir + "/" + base + ".ppm";
  std::FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return nullptr;
  int maxv = 0;
  if (std::fscanf(f, "P6 %d %d %d", w, h, &maxv) != 3 || maxv != 255) {
    std::fclose(f);
    return nullptr;
  }
  std::fgetc(f);
  g_images.emplace_back((std::size_t)*w * *h * 3u);
  const std::size_t got = std::fread(g_images.back().data(), 1, g_images.back().size(), f);
  std::fclose(f);
  return got == g_images.back().size() ? g_images.back().data() : nullptr;
}

void png_sink(const char*, int w, int h, int comp, const void* data, int) {
  if (g_png_path.empty()) return;
  std::FILE* f = std::fopen(g_png_path.c_str(), "wb");
  std::fwrite(data, 1, (std::size_t)w * h * comp, f);
  std::fclose(f);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <decoded-image-dir> <out.ptsc> [--run <out.rgb8>]\n", argv[0]);
    return 64;
  }
  g_image_dir = argv[1];
  g_out_path = argv[2];
  if (argc >= 5 && std::string(argv[3]) == "--run") {
    g_run = true;
    g_png_path = argv[4];
  }
  ptref_shim::image_provider() = provide;
  ptref_shim::png_sink() = png_sink;
  ptref_shim::hooks().on_host_buffer = on_buffer;
  ptref_shim::hooks().on_parallel_for = on_kernel;
  ptref_reference_main();
  return g_status;
}
