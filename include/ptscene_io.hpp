// ptscene_io.hpp -- PTSCENE1: a byte-for-byte serialisation of pt_scene + pt_camera
// (include/pt_abi.h).  Header-only C++17, no dependencies.
//
// Layout (little endian, natural alignment of the pt_abi.h PODs):
//   char     magic[8] = "PTSCENE1"
//   uint32_t n_hittables, n_spheres, n_rects, n_triangles, n_boxes, n_media,
//            n_materials, n_textures
//   uint64_t n_texture_bytes
//   int32_t  width, height, spp, depth      (the configuration the scene was built for)
//   pt_camera camera                        (96 bytes)
//   pt_order_entry[n_hittables] pt_sphere[] pt_rect[] pt_triangle[] pt_box[] pt_medium[]
//   pt_material[] pt_texture[] uint8_t texture_bytes[]
#ifndef PTSCENE_IO_HPP
#define PTSCENE_IO_HPP

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "pt_abi.h"

namespace ptscene {

struct file_meta {
  std::int32_t width, height, spp, depth;
};

// Owns the arrays a pt_scene points into.
struct owned_scene {
  std::vector<pt_order_entry> order;
  std::vector<pt_sphere> spheres;
  std::vector<pt_rect> rects;
  std::vector<pt_triangle> triangles;
  std::vector<pt_box> boxes;
  std::vector<pt_medium> media;
  std::vector<pt_material> materials;
  std::vector<pt_texture> textures;
  std::vector<std::uint8_t> texture_bytes;

  pt_scene view() const {
    pt_scene s {};
    s.n_hittables = (std::uint32_t)order.size(), s.order = order.data();
    s.n_spheres = (std::uint32_t)spheres.size(), s.spheres = spheres.data();
    s.n_rects = (std::uint32_t)rects.size(), s.rects = rects.data();
    s.n_triangles = (std::uint32_t)triangles.size(), s.triangles = triangles.data();
    s.n_boxes = (std::uint32_t)boxes.size(), s.boxes = boxes.data();
    s.n_media = (std::uint32_t)media.size(), s.media = media.data();
    s.n_materials = (std::uint32_t)materials.size(), s.materials = materials.data();
    s.n_textures = (std::uint32_t)textures.size(), s.textures = textures.data();
    s.n_texture_bytes = texture_bytes.size(), s.texture_bytes = texture_bytes.data();
    return s;
  }
};

namespace detail {
template <typename T> bool put(std::FILE* f, const std::vector<T>& v) {
  return v.empty() || std::fwrite(v.data(), sizeof(T), v.size(), f) == v.size();
}
template <typename T> bool get(std::FILE* f, std::vector<T>& v, std::size_t n) {
  v.resize(n);
  return n == 0 || std::fread(v.data(), sizeof(T), n, f) == n;
}
}  // namespace detail

inline bool save(const char* path, const owned_scene& s, const pt_camera& cam, const file_meta& meta) {
  std::FILE* f = std::fopen(path, "wb");
  if (!f) return false;
  const std::uint32_t counts[8] = { (std::uint32_t)s.order.size(),     (std::uint32_t)s.spheres.size(),
                                    (std::uint32_t)s.rects.size(),     (std::uint32_t)s.triangles.size(),
                                    (std::uint32_t)s.boxes.size(),     (std::uint32_t)s.media.size(),
                                    (std::uint32_t)s.materials.size(), (std::uint32_t)s.textures.size() };
  const std::uint64_t nbytes = s.texture_bytes.size();
  bool ok = std::fwrite("PTSCENE1", 1, 8, f) == 8 && std::fwrite(counts, 4, 8, f) == 8 &&
            std::fwrite(&nbytes, 8, 1, f) == 1 && std::fwrite(&meta, sizeof meta, 1, f) == 1 &&
            std::fwrite(&cam, sizeof cam, 1, f) == 1;
  ok = ok && detail::put(f, s.order) && detail::put(f, s.spheres) && detail::put(f, s.rects) &&
       detail::put(f, s.triangles) && detail::put(f, s.boxes) && detail::put(f, s.media) &&
       detail::put(f, s.materials) && detail::put(f, s.textures) && detail::put(f, s.texture_bytes);
  return std::fclose(f) == 0 && ok;
}

inline bool load(const char* path, owned_scene& s, pt_camera& cam, file_meta& meta) {
  std::FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  char magic[8];
  std::uint32_t counts[8];
  std::uint64_t nbytes = 0;
  bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "PTSCENE1", 8) == 0 &&
            std::fread(counts, 4, 8, f) == 8 && std::fread(&nbytes, 8, 1, f) == 1 &&
            std::fread(&meta, sizeof meta, 1, f) == 1 && std::fread(&cam, sizeof cam, 1, f) == 1;
  ok = ok && detail::get(f, s.order, counts[0]) && detail::get(f, s.spheres, counts[1]) &&
       detail::get(f, s.rects, counts[2]) && detail::get(f, s.triangles, counts[3]) &&
       detail::get(f, s.boxes, counts[4]) && detail::get(f, s.media, counts[5]) &&
       detail::get(f, s.materials, counts[6]) && detail::get(f, s.textures, counts[7]) &&
       detail::get(f, s.texture_bytes, (std::size_t)nbytes);
  std::fclose(f);
  return ok;
}

}  // namespace ptscene
#endif
