// TEST INFRASTRUCTURE ONLY.  Stand-in for the two stb_image entry points the
// reference calls (texture.hpp:103-109).  stb is an unpinned system package
// that is absent from this image, and file decode is outside the hot path:
// the oracle driver serves ALREADY-DECODED RGB8 texels through a provider
// callback, so the reference's own image_texture_factory copies them into its
// static pool unmodified.
#ifndef PT_ORACLE_STB_IMAGE_SHIM_H
#define PT_ORACLE_STB_IMAGE_SHIM_H
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace ptref_shim {
// Returns RGB8 texels (w*h*3 bytes, row 0 = top) for `name`, or nullptr.
// The returned memory stays owned by the provider.
using image_provider_t = const std::uint8_t* (*)(const char* name, int* w, int* h);
inline image_provider_t& image_provider() {
  static image_provider_t p = nullptr;
  return p;
}
}  // namespace ptref_shim

inline unsigned char* stbi_load(const char* name, int* w, int* h, int* comp, int req_comp) {
  if (req_comp != 3 || !ptref_shim::image_provider()) return nullptr;
  const std::uint8_t* src = ptref_shim::image_provider()(name, w, h);
  if (!src) return nullptr;
  if (comp) *comp = 3;
  const std::size_t n = static_cast<std::size_t>(*w) * static_cast<std::size_t>(*h) * 3u;
  auto* out = static_cast<unsigned char*>(std::malloc(n));
  std::memcpy(out, src, n);
  return out;  // the reference never frees it (texture.hpp:103-115)
}
inline const char* stbi_failure_reason() { return "pt oracle stb shim: no decoded image registered"; }
#endif
