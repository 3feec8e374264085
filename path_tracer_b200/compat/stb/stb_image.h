// Minimal stand-in for stb_image's stbi_load (image decode is host I/O outside the hot path; use the
// real stb_image.h ahead of this directory on the include path for jpg/png decoding).  Understands
// binary PPM (P6, maxval 255) and, for any other file name, a pre-decoded "<name>.ppm" beside it or
// in $PT_IMAGE_DIR/<basename>.ppm.
#ifndef PT_COMPAT_STB_IMAGE_H
#define PT_COMPAT_STB_IMAGE_H
#include <cstdio>
#include <cstdlib>
#include <string>

namespace pt_stb {
inline unsigned char* read_ppm(const std::string& path, int* w, int* h) {
  std::FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return nullptr;
  int maxv = 0;
  unsigned char* px = nullptr;
  if (std::fscanf(f, "P6 %d %d %d", w, h, &maxv) == 3 && maxv == 255 && *w > 0 && *h > 0) {
    std::fgetc(f);
    const std::size_t n = std::size_t(*w) * std::size_t(*h) * 3u;
    px = static_cast<unsigned char*>(std::malloc(n));
    if (std::fread(px, 1, n, f) != n) std::free(px), px = nullptr;
  }
  std::fclose(f);
  return px;
}
}  // namespace pt_stb

inline unsigned char* stbi_load(const char* name, int* w, int* h, int* comp, int req_comp) {
  if (req_comp != 3) return nullptr;
  if (comp) *comp = 3;
  const std::string path = name;
  if (unsigned char* p = pt_stb::read_ppm(path, w, h)) return p;
  if (unsigned char* p = pt_stb::read_ppm(path + ".ppm", w, h)) return p;
  if (const char* dir = std::getenv("PT_IMAGE_DIR")) {
    const auto slash = path.find_last_of('/');
    const std::string base = slash == std::string::npos ? path : path.substr(slash + 1);
    if (unsigned char* p = pt_stb::read_ppm(std::string(dir) + "/" + base + ".ppm", w, h)) return p;
  }
  return nullptr;
}
inline const char* stbi_failure_reason() { return "pt compat stb_image: only binary PPM (or a pre-decoded <file>.ppm) is understood"; }
inline void stbi_image_free(void* p) { std::free(p); }
#endif
