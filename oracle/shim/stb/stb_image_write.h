// TEST INFRASTRUCTURE ONLY.  Stand-in for stbi_write_png (reference
// src/main.cpp:57): hands the 8-bit image to an observer instead of encoding.
#ifndef PT_ORACLE_STB_IMAGE_WRITE_SHIM_H
#define PT_ORACLE_STB_IMAGE_WRITE_SHIM_H
namespace ptref_shim {
using png_sink_t = void (*)(const char* name, int w, int h, int comp, const void* data, int stride);
inline png_sink_t& png_sink() {
  static png_sink_t s = nullptr;
  return s;
}
}  // namespace ptref_shim
inline int stbi_write_png(const char* name, int w, int h, int comp, const void* data, int stride) {
  if (ptref_shim::png_sink()) ptref_shim::png_sink()(name, w, h, comp, data, stride);
  return 1;
}
#endif
