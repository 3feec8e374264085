// Minimal stand-in for stb_image_write's stbi_write_png: a valid PNG with stored (uncompressed)
// deflate blocks, no dependencies.  Use the real stb_image_write.h ahead of this directory on the
// include path for compressed output.
#ifndef PT_COMPAT_STB_IMAGE_WRITE_H
#define PT_COMPAT_STB_IMAGE_WRITE_H
#include <cstdint>
#include <cstdio>
#include <vector>

namespace pt_stb {
inline std::uint32_t crc32(const unsigned char* p, std::size_t n, std::uint32_t crc = 0) {
  static std::uint32_t table[256];
  if (!table[1])
    for (std::uint32_t i = 0; i < 256; ++i) {
      std::uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
  crc = ~crc;
  for (std::size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
  return ~crc;
}
inline void be32(std::vector<unsigned char>& v, std::uint32_t x) {
  for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s));
}
inline void chunk(std::FILE* f, const char* type, const std::vector<unsigned char>& data) {
  std::vector<unsigned char> buf;
  be32(buf, (std::uint32_t)data.size());
  buf.insert(buf.end(), type, type + 4);
  buf.insert(buf.end(), data.begin(), data.end());
  be32(buf, crc32(buf.data() + 4, buf.size() - 4));
  std::fwrite(buf.data(), 1, buf.size(), f);
}
}  // namespace pt_stb

inline int stbi_write_png(const char* name, int w, int h, int comp, const void* data, int stride) {
  if (comp != 3 && comp != 4) return 0;
  std::FILE* f = std::fopen(name, "wb");
  if (!f) return 0;
  const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
  std::fwrite(sig, 1, 8, f);
  std::vector<unsigned char> ihdr;
  pt_stb::be32(ihdr, (std::uint32_t)w), pt_stb::be32(ihdr, (std::uint32_t)h);
  ihdr.insert(ihdr.end(), { 8, (unsigned char)(comp == 3 ? 2 : 6), 0, 0, 0 });
  pt_stb::chunk(f, "IHDR", ihdr);
  // raw scanlines, each prefixed with filter type 0
  std::vector<unsigned char> raw;
  raw.reserve((std::size_t)h * ((std::size_t)w * comp + 1));
  for (int y = 0; y < h; ++y) {
    raw.push_back(0);
    const unsigned char* row = static_cast<const unsigned char*>(data) + (std::size_t)y * stride;
    raw.insert(raw.end(), row, row + (std::size_t)w * comp);
  }
  std::vector<unsigned char> z { 0x78, 0x01 };
  std::uint32_t a = 1, b = 0;
  for (unsigned char c : raw) a = (a + c) % 65521u, b = (b + a) % 65521u;
  for (std::size_t off = 0; off < raw.size() || off == 0; off += 65535) {
    const std::size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
    z.push_back(off + n >= raw.size() ? 1 : 0);
    z.push_back((unsigned char)(n & 0xff)), z.push_back((unsigned char)(n >> 8));
    z.push_back((unsigned char)(~n & 0xff)), z.push_back((unsigned char)((~n >> 8) & 0xff));
    z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    if (raw.empty()) break;
  }
  pt_stb::be32(z, (b << 16) | a);
  pt_stb::chunk(f, "IDAT", z);
  pt_stb::chunk(f, "IEND", {});
  return std::fclose(f) == 0;
}
#endif
