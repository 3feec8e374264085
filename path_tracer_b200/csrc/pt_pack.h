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

// What the chunk boxes are computed from (one per sphere ELEMENT, padding included).
struct SphereGeo {
  float c0[3], c1[3];
  float radius, time0, time1;
  bool valid;  // false: padding
};

struct PackedScene {
  std::vector<unsigned char> blob;
  uint32_t n_groups = 0;
  uint32_t off_groups = 0, off_sphere = 0, off_moving = 0, off_rect = 0, off_triangle = 0, off_box = 0;
  uint32_t off_trees = 0, off_nodes = 0, off_tree_ids = 0, n_trees = 0;
  uint32_t off_planes = 0, n_planes[3] = { 0, 0, 0 };  // sorted plane coordinates per axis (pt_packed.h: SceneDesc::off_planes)
  float flat_extent = 0.f;
  uint32_t off_sphere_box = 0, off_moving_box = 0;  // [kCullSets][chunks][2] float4 each, "no culling" until set
  std::vector<SphereGeo> sphere_geo, moving_geo;    // indexed like sphere_aux / moving_aux
  std::vector<unsigned char> sphere_chunk_open, moving_chunk_open;  // 1: chunk is never culled (outsized spheres)
  uint32_t n_objects = 0;
  uint32_t n_media_groups = 0, n_flat_groups = 0, n_late_sphere_groups = 0;
  std::vector<SphereAux> sphere_aux, moving_aux;
  std::vector<ObjAux> rect_aux, box_aux;
  std::vector<TriAux> tri_aux;
  std::vector<MediumRec> media;
  std::vector<pt_material> materials;
  std::vector<pt_texture> textures;
};

// Chunk bounding boxes for a camera whose rays carry times in [cam_time0, cam_time1] (camera.hpp:97-99):
// float4 {lo} {hi} per chunk, kCullSets sets per sphere kind, in the blob's layout.
struct CullBoxes {
  std::vector<float> sphere, moving;
  float bound[3];
};
void compute_cull_boxes(const PackedScene& ps, float cam_time0, float cam_time1, CullBoxes& out);

// The side tables both kernels index by scan id: the tie-break keys of every object (keys[key_base[type] + index]) and
// the way back, original object index -> scan id (-1: no such object).
void build_key_tables(const PackedScene& ps, std::vector<int32_t>& keys, uint32_t key_base[6], std::vector<int32_t>& object_id);

// Returns PT_OK or a PT_ERR_* code with a message in `error`.
int pack_scene(const pt_scene& scene, PackedScene& out, std::string& error);

}  // namespace ptb
#endif
