// pt_pack.h -- host side: flatten a pt_scene (include/pt_abi.h) into the device
// layout of pt_packed.h.  Replaces the reference's "wrap the variant vector in
// a sycl::buffer" step (render.hpp:146-148).
#ifndef PT_PACK_H
#define PT_PACK_H
#include <string>
#include <vector>

#include "pt_abi.h"
#include "pt_packed.h"

namespace ptb {

struct PackedScene {
  std::vector<unsigned char> blob;
  uint32_t n_groups = 0;
  uint32_t off_groups = 0, off_sphere = 0, off_moving = 0, off_rect = 0, off_triangle = 0, off_box = 0;
  uint32_t n_objects = 0;
  std::vector<SphereAux> sphere_aux, moving_aux;
  std::vector<ObjAux> rect_aux, box_aux;
  std::vector<TriAux> tri_aux;
  std::vector<MediumRec> media;
  std::vector<pt_material> materials;
  std::vector<pt_texture> textures;
};

// Returns PT_OK or a PT_ERR_* code with a message in `error`.
int pack_scene(const pt_scene& scene, PackedScene& out, std::string& error);

}  // namespace ptb
#endif
