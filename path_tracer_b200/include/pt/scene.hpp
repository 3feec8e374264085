// pt/scene.hpp -- host-side scene vocabulary with the names and constructor signatures that
// triSYCL/path_tracer scene scripts use (reference src/main.cpp), so such a
// script keeps compiling and builds the SAME scene, bit for bit.  These are plain value types:
// they only remember the constructor arguments (plus what the reference derives eagerly on the
// host, e.g. the clamped metal fuzz or the camera frame).  There are no hit()/scatter() methods
// here -- intersection and shading live in the CUDA kernel (csrc/pt_kernel.cu); the bridge is
// pt/flatten.hpp, which lowers a std::vector<hittable_t> to the C-ABI pt_scene.
//
// Reference counterparts (interfaces only): texture.hpp:18-154, material.hpp:11-135,
// sphere.hpp:26-49, rectangle.hpp:16-30, triangle.hpp:104-122, box.hpp:9-26,
// constant_medium.hpp:16-26, camera.hpp:67-87, rtweekend.hpp:21-57, render.hpp:22-23.
#ifndef PT_SCENE_HPP
#define PT_SCENE_HPP

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <limits>
#include <variant>
#include <vector>

#include "pt/sycl_facade.hpp"

#ifndef STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_IMPLEMENTATION
#endif
#include <stb/stb_image.h>

using real_t = float;
using point = sycl::float3;
using color = sycl::float3;
using vec = sycl::float3;

constexpr float infinity = std::numeric_limits<float>::infinity();
constexpr float pi = 3.1415926535897932385f;

inline float degrees_to_radians(float degrees) { return degrees * pi / 180.0f; }
inline vec unit_vector(const vec& v) { return v / sycl::length(v); }
inline vec operator-(const vec& v) { return vec(-v.x(), -v.y(), -v.z()); }

// ------------------------------------------------------------------ host random numbers
// Scene scripts randomise their scenes with this generator (main.cpp:76-99); the sequence has to
// be the reference's: 32-bit xorshift with the (7, 1, 9) triple, default state 2463534242.
class LocalPseudoRNG {
 public:
  explicit LocalPseudoRNG(std::uint32_t seed = 2463534242u) : state_ { seed } {}

  float float_t() {
    state_ ^= state_ >> 7;
    state_ ^= state_ << 1;
    state_ ^= state_ >> 9;
    return state_ * (1.f / 4294967296.f);
  }
  float float_t(float lo, float hi) { return lo + (hi - lo) * float_t(); }
  vec vec_t() {
    const float a = float_t();
    const float b = float_t();
    const float c = float_t();
    return vec(a, b, c);
  }
  vec vec_t(float lo, float hi) {
    const float span = hi - lo;
    return vec_t() * span + lo;
  }

 private:
  std::uint32_t state_;
};

// ------------------------------------------------------------------ textures
struct solid_texture {
  solid_texture() = default;
  solid_texture(const color& c) : rgb { c } {}
  solid_texture(float r, float g, float b) : rgb { r, g, b } {}
  color rgb;
};

struct checker_texture {
  checker_texture() = default;
  checker_texture(const solid_texture& first, const solid_texture& second) : odd { first }, even { second } {}
  checker_texture(const color& first, const color& second) : odd { first }, even { second } {}
  solid_texture odd, even;  // odd is used where sin(10x)sin(10y)sin(10z) < 0
};

// Image texels of every image_texture live in ONE process-wide RGB8 pool that starts with the
// fallback texel (0,0,1) used when a file cannot be loaded; a texture remembers its first texel.
// Decoding goes through the stb_image entry points (stbi_load / stbi_failure_reason), like the
// reference: put the real stb on the include path, or use the minimal stand-in shipped in
// path_tracer_b200/compat/stb (binary PPM, and "<file>.ppm" next to a jpg/png).

struct image_texture {
  std::size_t width = 1, height = 1, offset = 0;
  float cyclic_frequency = 1.f;

  static std::vector<std::uint8_t>& pool() {
    static std::vector<std::uint8_t> bytes { 0, 0, 1 };
    return bytes;
  }

  static image_texture image_texture_factory(const char* file_name, float cyclic_frequency = 1) {
    image_texture t;
    t.cyclic_frequency = cyclic_frequency;
    int w = 0, h = 0, comp = 3;
    unsigned char* texels = stbi_load(file_name, &w, &h, &comp, 3);
    if (!texels) {
      std::cerr << "ERROR: Could not load texture image file '" << file_name << "'.\n"
                << stbi_failure_reason() << std::endl;
      return t;  // 1x1 at texel 0
    }
    auto& bytes = pool();
    t.width = std::size_t(w), t.height = std::size_t(h), t.offset = bytes.size() / 3;
    bytes.insert(bytes.end(), texels, texels + std::size_t(3) * t.width * t.height);
    return t;
  }
};

using texture_t = std::variant<checker_texture, solid_texture, image_texture>;

// ------------------------------------------------------------------ materials
struct lambertian_material {
  lambertian_material() = default;
  lambertian_material(const color& a) : albedo { solid_texture { a } } {}
  lambertian_material(const texture_t& a) : albedo { a } {}
  texture_t albedo;
};

struct metal_material {
  metal_material() = default;
  metal_material(const color& a, float f) : albedo { a }, fuzz { std::clamp(f, 0.0f, 1.0f) } {}
  color albedo;
  float fuzz = 0.f;
};

struct dielectric_material {
  dielectric_material() = default;
  dielectric_material(real_t ri, const color& a) : ref_idx { ri }, albedo { a } {}
  real_t ref_idx = 1.f;
  color albedo;
};

struct lightsource_material {
  lightsource_material() = default;
  lightsource_material(const texture_t& a) : emit { a } {}
  lightsource_material(const color& a) : emit { solid_texture { a } } {}
  texture_t emit;
};

struct isotropic_material {
  isotropic_material(const color& a) : albedo { solid_texture { a } } {}
  isotropic_material(texture_t& a) : albedo { a } {}
  texture_t albedo;
};

using material_t = std::variant<lambertian_material, metal_material, dielectric_material, lightsource_material,
                                isotropic_material>;

// ------------------------------------------------------------------ hittables
struct sphere {
  sphere() = default;
  sphere(const point& c, real_t r, const material_t& m) : center0 { c }, center1 { c }, radius { r }, material_type { m } {}
  sphere(const point& c0, const point& c1, real_t t0, real_t t1, real_t r, const material_t& m)
      : center0 { c0 }, center1 { c1 }, radius { r }, time0 { t0 }, time1 { t1 }, material_type { m } {}
  point center0, center1;
  real_t radius = 0;
  real_t time0 = 0, time1 = 0;
  material_t material_type;
};

template <int Axis> struct axis_rect {  // Axis: 0 = xy (k is z), 1 = xz (k is y), 2 = yz (k is x)
  axis_rect() = default;
  axis_rect(real_t lo_a, real_t hi_a, real_t lo_b, real_t hi_b, real_t plane, const material_t& m)
      : a0 { lo_a }, a1 { hi_a }, b0 { lo_b }, b1 { hi_b }, k { plane }, material_type { m } {}
  real_t a0 = 0, a1 = 0, b0 = 0, b1 = 0, k = 0;
  material_t material_type;
};
using xy_rect = axis_rect<0>;
using xz_rect = axis_rect<1>;
using yz_rect = axis_rect<2>;
using rectangle_t = std::variant<xy_rect, xz_rect, yz_rect>;

struct triangle {
  triangle() = default;
  triangle(const point& a, const point& b, const point& c, const material_t& m) : v0 { a }, v1 { b }, v2 { c }, material_type { m } {}
  point v0, v1, v2;
  material_t material_type;
};

struct box {
  box() = default;
  box(const point& lo, const point& hi, const material_t& m) : box_min { lo }, box_max { hi }, material_type { m } {}
  point box_min, box_max;
  material_t material_type;
};

using hittableVolume_t = std::variant<sphere, box>;

struct constant_medium {
  constant_medium(const hittableVolume_t& b, real_t d, texture_t& a) : boundary { b }, density { d }, phase_function { isotropic_material { a } } {}
  constant_medium(const hittableVolume_t& b, real_t d, const color& a) : boundary { b }, density { d }, phase_function { isotropic_material { a } } {}
  hittableVolume_t boundary;
  real_t density;  // the kernel uses -1/density, derived when the scene is uploaded
  material_t phase_function;
};

using hittable_t = std::variant<sphere, xy_rect, triangle, box, constant_medium>;

// ------------------------------------------------------------------ camera
// Thin-lens camera.  The constructor derives the frame exactly like the reference does
// (camera.hpp:67-87: same operations, same order), because those 24 floats are kernel input.
class camera {
 public:
  camera(const point& look_from, const point& look_at, const vec& vup, real_t degree_vfov, real_t aspect_ratio,
         real_t aperture, real_t focus_dist, real_t shutter_open = 0, real_t shutter_close = 0) {
    const real_t half_height = sycl::tan(degrees_to_radians(degree_vfov) / 2);
    const real_t viewport_height = 2.0f * half_height;
    const real_t viewport_width = aspect_ratio * viewport_height;
    origin = look_from;
    w = unit_vector(look_from - look_at);
    u = unit_vector(sycl::cross(vup, w));
    v = sycl::cross(w, u);
    horizontal = focus_dist * viewport_width * u;
    vertical = focus_dist * viewport_height * v;
    lower_left_corner = origin - horizontal / 2 - vertical / 2 - focus_dist * w;
    lens_radius = aperture / 2;
    time0 = shutter_open;
    time1 = shutter_close;
  }
  // field order = pt_camera (include/pt_abi.h)
  point origin, lower_left_corner;
  vec horizontal, vertical, u, v, w;
  real_t lens_radius, time0, time1;
};
static_assert(sizeof(camera) == 96, "camera must mirror pt_camera");

#endif
